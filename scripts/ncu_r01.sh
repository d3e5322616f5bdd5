# usage: bash scripts/ncu_r01.sh   (on a GPU box; writes gpurun_out/r01b_*)  -- launch list + full capture of the top kernels
mkdir -p gpurun_out
CMD="python bench.py --seqs 64 --batches 1 --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 260 --csv --log-file gpurun_out/r01b_launches.csv $CMD > gpurun_out/r01b_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lk_kernel_fast|pyrdown_kernel|imu_cov_propagate|fast_kernel|ekf_gain|ekf_cov" -s 70 -c 12 -o gpurun_out/r01b_top -f $CMD > gpurun_out/r01b_top.log 2>&1
ls -la gpurun_out/ | head -20
