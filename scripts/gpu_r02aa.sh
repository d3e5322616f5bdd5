# round 2: LK at 8 CTAs per SM (64 registers) against 6 (80 registers)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tracker.py -x -q -m gpu -k "lk" > gpurun_out/r02aa_pytest6.txt 2>&1; tail -1 gpurun_out/r02aa_pytest6.txt
XIVO_LK_MINB=8 timeout 600 python -m pytest tests/test_gpu_tracker.py -x -q -m gpu -k "lk" > gpurun_out/r02aa_pytest8.txt 2>&1; tail -1 gpurun_out/r02aa_pytest8.txt
run() {
  name=$1; shift
  timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-single-stream "$@" > gpurun_out/r02aa_$name.json 2> gpurun_out/r02aa_$name.err
  python - "$name" <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02aa_{n}.json').read().strip().splitlines()[-1])
    print(n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'seqs',d['config']['sequences_per_gpu'])
    ks=d['roofline']['kernels']
    print('  kernels us/launch:', {k: round(v['ms']*1000/max(v['calls'],1),1) for k,v in ks.items()})
except Exception as e:
    print(n,'FAILED',e); print(open(f'gpurun_out/r02aa_{n}.err').read()[-1500:])
P
}
run lk6a
XIVO_LK_MINB=8 run lk8a
run lk6b
XIVO_LK_MINB=8 run lk8b
