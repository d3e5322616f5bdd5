# round 2: frame prefetch (identity test, e2e with / without), 1024 x 32 default, driver-style run
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_estimator.py -x -q -m gpu -k "prefetch or step_call" > gpurun_out/r02x_pytest.txt 2>&1
tail -3 gpurun_out/r02x_pytest.txt
run() {
  name=$1; shift
  timeout 500 python bench.py "$@" > gpurun_out/r02x_$name.json 2> gpurun_out/r02x_$name.err
  python - "$name" <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02x_{n}.json').read().strip().splitlines()[-1])
    print(n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'seqs',d['config']['sequences_per_gpu'],'batches',d['config']['batches_per_gpu'],'prefetch',d['config'].get('frame_prefetch'),'cpu',(d.get('cpu_baseline') or {}).get('value'))
except Exception as e:
    print(n,'FAILED',e); print(open(f'gpurun_out/r02x_{n}.err').read()[-1500:])
P
}
run default --gpus 1 --steps 20 --warmup 5
run nopf --steps 10 --warmup 3 --no-cpu-baseline --no-single-stream --no-prefetch
run pf --steps 10 --warmup 3 --no-cpu-baseline --no-single-stream
run hostprof_e2e --steps 10 --warmup 3 --no-cpu-baseline --no-single-stream --profile-e2e --profile-overlapped --profile-level 3
