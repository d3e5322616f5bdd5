# round 2: prefetch with the per-slot ingest wait, interleaved with the plain path
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_estimator.py -x -q -m gpu -k "prefetch or step_call" > gpurun_out/r02y_pytest.txt 2>&1
tail -2 gpurun_out/r02y_pytest.txt
run() {
  name=$1; shift
  timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-single-stream "$@" > gpurun_out/r02y_$name.json 2> gpurun_out/r02y_$name.err
  python - "$name" <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02y_{n}.json').read().strip().splitlines()[-1])
    print(n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'seqs',d['config']['sequences_per_gpu'],'prefetch',d['config'].get('frame_prefetch'))
    hp=d.get('host_phase_ms_per_frame_step') or {}
    nb=d['config']['batches_per_gpu']
    if 'prof' in n:
        print('  per batch-frame ms:', {k: round(v/nb,3) for k,v in sorted(hp.items()) if not k.startswith('x_')})
except Exception as e:
    print(n,'FAILED',e); print(open(f'gpurun_out/r02y_{n}.err').read()[-1500:])
P
}
run pf1
run nopf1 --no-prefetch
run pf2
run nopf2 --no-prefetch
run pfprof --profile-e2e --profile-overlapped --profile-level 3
