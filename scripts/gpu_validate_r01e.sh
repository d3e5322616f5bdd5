# usage: bash scripts/gpu_validate_r01e.sh (GPU box) -- last call of round 1: second IMU-covariance formulation + narrower tcgen05 tiles,
# the vio app test, final bench line with CPU baseline, A/B against the first IMU formulation, ncu of the changed kernels
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_ekf.py tests/test_gpu_estimator.py tests/test_cpp_io.py tests/test_cpp_facade.py -m gpu -q -rf --timeout 120 -p no:cacheprovider > gpurun_out/r01e_pytest_default.txt 2>&1
tail -6 gpurun_out/r01e_pytest_default.txt
XIVO_IMU_V1=1 timeout 120 python -m pytest tests/test_gpu_ekf.py tests/test_gpu_estimator.py -m gpu -q -rf --timeout 120 -p no:cacheprovider -k "imu or pcw_trajectory_parity or image_pipeline or step_call" > gpurun_out/r01e_pytest_imu_v1.txt 2>&1
tail -3 gpurun_out/r01e_pytest_imu_v1.txt
timeout 240 python bench.py > gpurun_out/r01e_bench.json 2> gpurun_out/r01e_bench.err
head -c 300 gpurun_out/r01e_bench.json; echo
XIVO_IMU_V1=1 timeout 120 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r01e_bench_imu_v1.json 2> gpurun_out/r01e_bench_imu_v1.err
head -c 300 gpurun_out/r01e_bench_imu_v1.json; echo
CMD="python bench.py --seqs 64 --batches 1 --steps 2 --warmup 3 --no-cpu-baseline --cov-update tf32x3"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"imu_cov_propagate|ekf_cov_tc" -s 30 -c 4 -o gpurun_out/r01e_top -f $CMD > gpurun_out/r01e_top.log 2>&1
tail -2 gpurun_out/r01e_top.log
