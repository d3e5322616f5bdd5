# round 2, GPU call 8: CUDA-call diet (ticket waits, one upload per phase, edit lists folded into the gate / gain kernels), TMA pyramid + FAST by default
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf --timeout 600 -p no:cacheprovider -x > gpurun_out/r02h_pytest_gpu.txt 2>&1
tail -15 gpurun_out/r02h_pytest_gpu.txt
run() {
  name=$1; shift
  envs=""; while [ "$1" != "--" ]; do envs="$envs $1"; shift; done; shift
  env $envs timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-single-stream "$@" > gpurun_out/r02h_$name.json 2> gpurun_out/r02h_$name.err
  python - "$name" <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02h_{n}.json').read().strip().splitlines()[-1])
    print(n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'ingest',d['config'].get('frame_ingest'),d['config'].get('frame_ingest_calibration_ms_per_step'),'launches',d['gpu_launches'])
    hp=d.get('host_phase_ms_per_frame_step') or {}
    nb=d['config']['batches_per_gpu']
    if hp and 'prof' in n: print('  per batch-frame ms:', {k: round(v/nb,3) for k,v in sorted(hp.items()) if not k.startswith('x_')})
    if 'prof' not in n: print('  kernels us/launch:', {k: round(v['ms']*1e3/v['calls'],1) for k,v in d['roofline']['kernels'].items()})
except Exception as e:
    print(n,'FAILED',e); print(open(f'gpurun_out/r02h_{n}.err').read()[-1500:])
P
}
run default --
run prof -- --profile-overlapped
run prof_e2e -- --profile-overlapped --profile-e2e
run notma XIVO_PYRDOWN_TMA=0 XIVO_FAST_TMA=0 --
run nomemops XIVO_NO_STREAM_MEMOPS=1 --
run b16 -- --batches 16
