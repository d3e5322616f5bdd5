# usage: bash scripts/gpu_validate_r02.sh (GPU box) -- FIRST call of round 2.  Everything committed after the last GPU minute of round 1
# was verified on the CPU only (tests/test_host_twin.py, test_reference_pin.py, test_host_logic.py); this confirms it on the B200:
#   1. the whole GPU suite (host orders of libstdc++, triangulate_pre_subfilter, the complete read-back surface, tracker list order)
#   2. the bench line with the frame-ingest calibration (zero-copy gather vs copy engine), plus each mode forced, for the A/B
#   3. launch list + one ncu --set full capture of the two top kernels of the same command
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rf --timeout 300 -p no:cacheprovider > gpurun_out/r02a_pytest_gpu.txt 2>&1
tail -8 gpurun_out/r02a_pytest_gpu.txt
timeout 300 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
head -c 400 gpurun_out/r02a_bench.json; echo; grep "frame ingest" gpurun_out/r02a_bench.err
for m in zero_copy copy_engine; do
  timeout 150 python bench.py --steps 20 --no-cpu-baseline --ingest $m > gpurun_out/r02a_bench_$m.json 2> gpurun_out/r02a_bench_$m.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r02a_bench_$m.json').read().strip().splitlines()[-1]); print('$m', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['e2e']['ms_per_step'],3))"
done
timeout 300 python scripts/explore_image_options.py > gpurun_out/r02a_explore.txt 2>&1; tail -8 gpurun_out/r02a_explore.txt
CMD="python bench.py --seqs 64 --batches 1 --steps 2 --warmup 3 --no-cpu-baseline --ingest zero_copy"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02a_launches.csv $CMD > gpurun_out/r02a_launches.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"lk_kernel|pyrdown" -s 40 -c 4 -o gpurun_out/r02a_top -f $CMD > gpurun_out/r02a_top.log 2>&1
tail -2 gpurun_out/r02a_top.log
