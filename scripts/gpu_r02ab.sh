# round 2: more batches in flight at 1024 / 1536 sequences
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-single-stream "$@" > gpurun_out/r02ab_$name.json 2> gpurun_out/r02ab_$name.err
  python - "$name" <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02ab_{n}.json').read().strip().splitlines()[-1])
    print(n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'seqs',d['config']['sequences_per_gpu'],'batches',d['config']['batches_per_gpu'],'threads',d['config'].get('host_threads'))
except Exception as e:
    print(n,'FAILED',e); print(open(f'gpurun_out/r02ab_{n}.err').read()[-1500:])
P
}
run b8
run b12 --batches 12
run b10 --batches 10
run s1536b12 --seqs 1536 --streams 48 --batches 12
