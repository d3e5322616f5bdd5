# round 2, GPU call 7: CUDA API contention probe; where the driver threads spend their time (issue_* scopes); batch-split sweep with copied tables
mkdir -p gpurun_out
timeout 120 scripts/probes/api_probe 2>&1 | tee gpurun_out/r02g_api_probe.txt
run() {
  name=$1; shift
  envs=""; while [ "$1" != "--" ]; do envs="$envs $1"; shift; done; shift
  env $envs timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-single-stream "$@" > gpurun_out/r02g_$name.json 2> gpurun_out/r02g_$name.err
  python - "$name" <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02g_{n}.json').read().strip().splitlines()[-1])
    print(n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'ingest',d['config'].get('frame_ingest'),d['config'].get('frame_ingest_calibration_ms_per_step'))
    hp=d.get('host_phase_ms_per_frame_step') or {}
    nb=d['config']['batches_per_gpu']
    if hp: print('  per batch-frame ms:', {k: round(v/nb,3) for k,v in sorted(hp.items()) if not k.startswith('x_')})
except Exception as e:
    print(n,'FAILED',e); print(open(f'gpurun_out/r02g_{n}.err').read()[-800:])
P
}
run pool8_prof -- --batches 8 --profile-overlapped
run pool8_prof_e2e -- --batches 8 --profile-overlapped --profile-e2e
run pool4 -- --batches 4
run pool2 -- --batches 2
run pool8_nohelp XIVO_HELP=0 -- --batches 8
run pool8_zc_ingest -- --batches 8 --ingest zero_copy
