# round 2, GPU call 2: lanes (one library thread per sub-batch), device-side tracker decisions, pitched H2D copies, the new bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -rf --timeout 600 -p no:cacheprovider > gpurun_out/r02b_pytest_gpu.txt 2>&1
tail -15 gpurun_out/r02b_pytest_gpu.txt
run() {  # name, env..., -- bench args
  name=$1; shift
  envs=""; while [ "$1" != "--" ]; do envs="$envs $1"; shift; done; shift
  env $envs timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-single-stream "$@" > gpurun_out/r02b_$name.json 2> gpurun_out/r02b_$name.err
  python - "$name" <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02b_{n}.json').read().strip().splitlines()[-1])
    print(n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'lanes',d['config'].get('lanes_per_batch'),'ingest',d['config'].get('frame_ingest'),d['config'].get('frame_ingest_calibration_ms_per_step'),'tracked',round(d['tracked_features_mean']), 'busy', round(d['roofline']['device_busy_frac'],2))
except Exception as e:
    print(n,'FAILED',e); print(open(f'gpurun_out/r02b_{n}.err').read()[-1500:])
P
}
run lanes_default -- 
run lanes8 XIVO_LANES=8 --
run lanes16 XIVO_LANES=16 --
run lanes32 XIVO_LANES=32 --
run lanes48 XIVO_LANES=48 --
run lanes32_tok12 XIVO_LANES=32 XIVO_CPU_TOKENS=12 --
run lanes32_tok20 XIVO_LANES=32 XIVO_CPU_TOKENS=20 --
run lanes_default_conn8 CUDA_DEVICE_MAX_CONNECTIONS=8 --
run lanes22_hostdec XIVO_HOST_TRACKER_DECISIONS=1 --
run legacy_pool XIVO_LANES=1 -- --batches 8
run legacy_pool_hostdec XIVO_LANES=1 XIVO_HOST_TRACKER_DECISIONS=1 -- --batches 8
timeout 300 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r02b_bench_full.json 2> gpurun_out/r02b_bench_full.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02b_bench_full.json').read().strip().splitlines()[-1])
print('FULL value',round(d['value']),'e2e',round(d['e2e']['value']),'single',d.get('single_stream'))
print({k:(round(v['ms'],2),v['calls']) for k,v in d['roofline']['kernels'].items()})
print(d['host_phase_ms_per_frame_step'])
P
