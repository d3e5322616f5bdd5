# round 2: full GPU suite + bench with the lazy-adjacency host code; batch-split sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf --timeout 600 -p no:cacheprovider -x > gpurun_out/r02r_pytest_gpu.txt 2>&1
tail -4 gpurun_out/r02r_pytest_gpu.txt
run() {
  name=$1; shift
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-single-stream "$@" > gpurun_out/r02r_$name.json 2> gpurun_out/r02r_$name.err
  python - "$name" <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02r_{n}.json').read().strip().splitlines()[-1])
    print(n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'ingest',d['config'].get('frame_ingest'))
    hp=d.get('host_phase_ms_per_frame_step') or {}
    if 'prof' in n: print('  per sequence-frame us:', {k: round(v*1000/d['config']['sequences_per_gpu'],2) for k,v in sorted(hp.items()) if k.startswith('x_') and not k.startswith('x_i_')})
except Exception as e:
    print(n,'FAILED',e); print(open(f'gpurun_out/r02r_{n}.err').read()[-1200:])
P
}
run default --
run b6 -- --batches 6
run b12 -- --batches 12
run hostprof -- --profile-overlapped --profile-level 3
