# round 2: bench after the host-side reductions (lazy adjacency, Feature layout, hoisted RK invariants, parallel stage staging); batch-split sweep; host-only profile
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-single-stream "$@" > gpurun_out/r02s_$name.json 2> gpurun_out/r02s_$name.err
  python - "$name" <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02s_{n}.json').read().strip().splitlines()[-1])
    print(n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'ingest',d['config'].get('frame_ingest'))
    hp=d.get('host_phase_ms_per_frame_step') or {}
    nb=d['config']['batches_per_gpu']
    if 'prof' in n:
        print('  per batch-frame ms:', {k: round(v/nb,3) for k,v in sorted(hp.items()) if not k.startswith('x_')})
        print('  per sequence-frame us:', {k: round(v*1000/d['config']['sequences_per_gpu'],2) for k,v in sorted(hp.items()) if k.startswith('x_') and not k.startswith('x_i_')})
except Exception as e:
    print(n,'FAILED',e); print(open(f'gpurun_out/r02s_{n}.err').read()[-1200:])
P
}
run default
run b6 --batches 6
run b12 --batches 12
run hostprof --profile-overlapped --profile-level 3
