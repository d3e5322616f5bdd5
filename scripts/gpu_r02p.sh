# round 2: bench lines of BASELINE configs[2] / [3] (fp64 and tensor-core downdate), launch list + full ncu of the default bench's kernels with the final code
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-single-stream "$@" > gpurun_out/r02p_$name.json 2> gpurun_out/r02p_$name.err
  python - "$name" <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02p_{n}.json').read().strip().splitlines()[-1])
    r=d['roofline']
    print(n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'tracked',round(d['tracked_features_mean']),'ekf_update',r.get('ekf_update'))
    print('  kernels us/launch:', {k: round(v['ms']*1e3/v['calls'],1) for k,v in r['kernels'].items()})
except Exception as e:
    print(n,'FAILED',e); print(open(f'gpurun_out/r02p_{n}.err').read()[-1200:])
P
}
run config2 --config 2
run config2_tc --config 2 --cov-update tf32x3
run config3 --config 3
run config3_tc --config 3 --cov-update tf32x3
CMD="python bench.py --seqs 64 --batches 1 --steps 1 --warmup 3 --frames-per-step 2 --no-cpu-baseline --no-single-stream --ingest copy_engine"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02p_launches.csv $CMD > gpurun_out/r02p_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"lk_kernel|pyrdown_tma|fast_pair_tma|track_|imu_cov|ekf_gain|ekf_cov|jacobian|subfilter|fetch" -s 500 -c 24 -o gpurun_out/r02p_top -f $CMD > gpurun_out/r02p_top.log 2>&1
tail -2 gpurun_out/r02p_top.log
