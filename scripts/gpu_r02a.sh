# round 2, GPU call 1: validate everything committed after the last GPU minute of round 1 (+ this round's fixes) and collect the
# baseline the round starts from: full GPU suite, bench line, fine host-phase profile, ingest A/B, PCIe ceilings, launch list + ncu.
mkdir -p gpurun_out
nproc; cat /sys/fs/cgroup/cpu.max; nvidia-smi --query-gpu=name,pcie.link.gen.max,pcie.link.width.max --format=csv
timeout 1200 python -m pytest tests -m gpu -q -rf --timeout 600 -p no:cacheprovider --durations=15 > gpurun_out/r02a_pytest_gpu.txt 2>&1
tail -25 gpurun_out/r02a_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 120 python scripts/pcie_probe.py > gpurun_out/r02a_pcie.txt 2>&1; cat gpurun_out/r02a_pcie.txt
timeout 300 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02a_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'cpu',d['cpu_baseline'] and round(d['cpu_baseline']['value']), d['config'].get('frame_ingest'), d['config'].get('frame_ingest_calibration_ms_per_step'))
print({k:(round(v['ms'],2),v['calls']) for k,v in d['roofline']['kernels'].items()}, d['roofline']['device_busy_frac'])
print(d['host_phase_ms_per_step'])
P
timeout 200 python bench.py --steps 20 --no-cpu-baseline --profile-level 2 > gpurun_out/r02a_bench_fine.json 2> gpurun_out/r02a_bench_fine.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02a_bench_fine.json').read().strip().splitlines()[-1])
print('fine', d['host_phase_ms_per_step'], 'value', round(d['value']))
P
grep -i "fine\|x_" gpurun_out/r02a_bench_fine.err | tail -5
CMD="python bench.py --seqs 64 --batches 1 --steps 2 --warmup 3 --no-cpu-baseline --ingest zero_copy"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02a_launches.csv $CMD > gpurun_out/r02a_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"lk_kernel|pyrdown|fast_|imu_cov|ekf_gain|subfilter|cov_edit" -s 60 -c 14 -o gpurun_out/r02a_top -f $CMD > gpurun_out/r02a_top.log 2>&1
tail -2 gpurun_out/r02a_top.log; ls -la gpurun_out | tail -12
