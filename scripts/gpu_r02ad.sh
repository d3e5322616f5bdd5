# round 2, final code: whole GPU suite, smoke, driver-style bench lines, launch list + full ncu of the kernels of one 128-sequence batch
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02ad_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r02ad_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
run() {
  name=$1; shift
  timeout 600 python bench.py "$@" > gpurun_out/r02ad_$name.json 2> gpurun_out/r02ad_$name.err
  python - "$name" <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02ad_{n}.json').read().strip().splitlines()[-1])
    print(n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'cpu',(d.get('cpu_baseline') or {}).get('value'),'threads',d['config'].get('host_threads'),'cpus_used',(d.get('host_cpu') or {}).get('value_pass'))
except Exception as e:
    print(n,'FAILED',e); print(open(f'gpurun_out/r02ad_{n}.err').read()[-1500:])
P
}
run default --gpus 1 --steps 20 --warmup 5
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02ad_reference.json 2> gpurun_out/r02ad_reference.err; tail -c 600 gpurun_out/r02ad_reference.json
run h0 --steps 10 --warmup 3 --no-cpu-baseline --no-single-stream --cpu-headroom 0
CMD="python bench.py --seqs 128 --streams 4 --batches 1 --steps 1 --warmup 3 --frames-per-step 2 --no-cpu-baseline --no-single-stream --ingest copy_engine"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02ad_launches.csv $CMD > gpurun_out/r02ad_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"lk_kernel|pyrdown_tma|fast_pair_tma|track_|imu_cov|ekf_gain|ekf_cov|jacobian|subfilter|fetch" -s 520 -c 22 -o gpurun_out/r02ad_top -f $CMD > gpurun_out/r02ad_top.log 2>&1
tail -2 gpurun_out/r02ad_top.log
