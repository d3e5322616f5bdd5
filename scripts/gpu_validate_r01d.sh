# usage: bash scripts/gpu_validate_r01d.sh  (GPU box) -- bring-up of the tcgen05 downdate + the whole GPU suite + tensor-mode bench
mkdir -p gpurun_out
SW=0
for sw in 0 1; do
  XIVO_TC_SWAP=$sw timeout 90 python scripts/tc_probe.py > gpurun_out/r01d_tc_probe_$sw.txt 2>&1 && SW=$sw && break
done
tail -3 gpurun_out/r01d_tc_probe_*.txt
echo "descriptor variant used for the tests: XIVO_TC_SWAP=$SW" | tee gpurun_out/r01d_variant.txt
export XIVO_TC_SWAP=$SW
timeout 420 python -m pytest tests/test_gpu_ekf.py tests/test_gpu_estimator.py tests/test_cpp_facade.py tests/test_gpu_tracker.py tests/test_replicas_gloo.py -m gpu -q -rf --timeout 150 -p no:cacheprovider --durations=8 > gpurun_out/r01d_pytest.txt 2>&1
tail -25 gpurun_out/r01d_pytest.txt
timeout 90 ncu --set full --clock-control none --import-source on -k regex:ekf_cov_tc -c 3 -o gpurun_out/r01d_tc -f python scripts/tc_probe.py > gpurun_out/r01d_tc_ncu.log 2>&1
timeout 150 python bench.py --cov-update tf32x3 --steps 20 --no-cpu-baseline > gpurun_out/r01d_bench_tf32x3.json 2> gpurun_out/r01d_bench_tf32x3.err
tail -2 gpurun_out/r01d_bench_tf32x3.err
head -c 600 gpurun_out/r01d_bench_tf32x3.json
