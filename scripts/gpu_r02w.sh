# round 2: pairwise-coverage track_accept; 1024 sequences x 32 base streams against 512 x 16, repeated to see the run-to-run spread
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tracker_decisions.py tests/test_gpu_tracker.py tests/test_gpu_bench_workload.py -x -q -m gpu > gpurun_out/r02w_pytest.txt 2>&1
tail -2 gpurun_out/r02w_pytest.txt
run() {
  name=$1; shift
  timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-single-stream "$@" > gpurun_out/r02w_$name.json 2> gpurun_out/r02w_$name.err
  python - "$name" <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02w_{n}.json').read().strip().splitlines()[-1])
    print(n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'seqs',d['config']['sequences_per_gpu'],'batches',d['config']['batches_per_gpu'])
    ks=d['roofline']['kernels']
    print('  kernels us/launch:', {k: round(v['ms']*1000/max(v['calls'],1),1) for k,v in ks.items()})
except Exception as e:
    print(n,'FAILED',e); print(open(f'gpurun_out/r02w_{n}.err').read()[-1200:])
P
}
run s1024a --seqs 1024 --streams 32
run s512a --seqs 512 --streams 16
run s1024b --seqs 1024 --streams 32
run s512b --seqs 512 --streams 16
