# round 2, GPU call 9: host-only profile (no kernel events) of the driver threads, device-resident and e2e
mkdir -p gpurun_out
run() {
  name=$1; shift
  envs=""; while [ "$1" != "--" ]; do envs="$envs $1"; shift; done; shift
  env $envs timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-single-stream "$@" > gpurun_out/r02i_$name.json 2> gpurun_out/r02i_$name.err
  python - "$name" <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02i_{n}.json').read().strip().splitlines()[-1])
    print(n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'ingest',d['config'].get('frame_ingest'),d['config'].get('frame_ingest_calibration_ms_per_step'),'prof ms/step',d['roofline'].get('profiled_pass_ms_per_step'))
    hp=d.get('host_phase_ms_per_frame_step') or {}
    nb=d['config']['batches_per_gpu']
    print('  per batch-frame ms (driver scopes):', {k: round(v/nb,3) for k,v in sorted(hp.items()) if not k.startswith('x_') or k.startswith('x_i_')})
    print('  per sequence-frame us:', {k: round(v*1000/d['config']['sequences_per_gpu'],2) for k,v in sorted(hp.items()) if k.startswith('x_') and not k.startswith('x_i_')})
except Exception as e:
    print(n,'FAILED',e); print(open(f'gpurun_out/r02i_{n}.err').read()[-1500:])
P
}
run hostprof -- --profile-overlapped --profile-level 3
run hostprof_e2e -- --profile-overlapped --profile-level 3 --profile-e2e
run hostprof_e2e_ce -- --profile-overlapped --profile-level 3 --profile-e2e --ingest copy_engine
