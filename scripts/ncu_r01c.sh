# usage: bash scripts/ncu_r01c.sh   (on a GPU box; writes gpurun_out/r01c_*)
# default bench line, then launch list + full capture of the top kernels (numbers under ncu are never bench values)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r01c_env.txt 2>&1
(nproc; cat /sys/fs/cgroup/cpu.max; lscpu | head -20) >> gpurun_out/r01c_env.txt 2>&1
timeout 480 python bench.py > gpurun_out/r01c_bench.json 2> gpurun_out/r01c_bench.err
tail -3 gpurun_out/r01c_bench.err
CMD="python bench.py --seqs 64 --batches 1 --steps 2 --warmup 3 --no-cpu-baseline"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 260 --csv --log-file gpurun_out/r01c_launches.csv $CMD > gpurun_out/r01c_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"lk_kernel_fast|pyrdown|imu_cov_propagate|fast_kernel|fast_pair|ekf_gain|ekf_cov" -s 70 -c 12 -o gpurun_out/r01c_top -f $CMD > gpurun_out/r01c_top.log 2>&1
ls -la gpurun_out/ | head -20
head -c 1500 gpurun_out/r01c_bench.json
