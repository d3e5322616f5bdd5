# round 2: prefetch queue (depth 2) + NUMA placement: identity test, e2e with / without each, interleaved
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_estimator.py -x -q -m gpu -k "prefetch or step_call" > gpurun_out/r02z_pytest.txt 2>&1
tail -4 gpurun_out/r02z_pytest.txt
run() {
  name=$1; shift
  timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-single-stream "$@" > gpurun_out/r02z_$name.json 2> gpurun_out/r02z_$name.err
  python - "$name" <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02z_{n}.json').read().strip().splitlines()[-1])
    print(n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'seqs',d['config']['sequences_per_gpu'],'prefetch',d['config'].get('frame_prefetch'),'numa',d['config'].get('numa'))
    hp=d.get('host_phase_ms_per_frame_step') or {}
    nb=d['config']['batches_per_gpu']
    if 'prof' in n:
        print('  per batch-frame ms:', {k: round(v/nb,3) for k,v in sorted(hp.items()) if not k.startswith('x_')})
except Exception as e:
    print(n,'FAILED',e); print(open(f'gpurun_out/r02z_{n}.err').read()[-1500:])
P
}
lscpu | grep -i "numa" > gpurun_out/r02z_numa.txt; nvidia-smi topo -m >> gpurun_out/r02z_numa.txt 2>&1
run pf_numa1
run nopf_nonuma1 --no-prefetch --no-numa
run pf_nonuma --no-numa
run nopf_numa --no-prefetch
run pf_numa2
run nopf_nonuma2 --no-prefetch --no-numa
run pfprof --profile-e2e --profile-overlapped --profile-level 3
