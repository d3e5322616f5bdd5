# round 2, GPU call 10: table uploads by fetch kernels (mapped pinned -> device), frames by the copy engine
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf --timeout 600 -p no:cacheprovider -x > gpurun_out/r02j_pytest_gpu.txt 2>&1
tail -5 gpurun_out/r02j_pytest_gpu.txt
run() {
  name=$1; shift
  envs=""; while [ "$1" != "--" ]; do envs="$envs $1"; shift; done; shift
  env $envs timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-single-stream "$@" > gpurun_out/r02j_$name.json 2> gpurun_out/r02j_$name.err
  python - "$name" <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02j_{n}.json').read().strip().splitlines()[-1])
    print(n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'ingest',d['config'].get('frame_ingest'),d['config'].get('frame_ingest_calibration_ms_per_step'),'prof ms/step',d['roofline'].get('profiled_pass_ms_per_step'))
    hp=d.get('host_phase_ms_per_frame_step') or {}
    nb=d['config']['batches_per_gpu']
    if 'prof' in n: print('  per batch-frame ms (driver scopes):', {k: round(v/nb,3) for k,v in sorted(hp.items()) if not k.startswith('x_')})
except Exception as e:
    print(n,'FAILED',e); print(open(f'gpurun_out/r02j_{n}.err').read()[-1500:])
P
}
run default --
run hostprof_e2e -- --profile-overlapped --profile-level 3 --profile-e2e
run b16 -- --batches 16
