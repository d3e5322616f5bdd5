# round 2: staged track_accept + feature prefetch: decision parity tests, driver-style bench, e2e host profile
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tracker_decisions.py tests/test_gpu_tracker.py tests/test_gpu_bench_workload.py -x -q -m gpu > gpurun_out/r02u_pytest.txt 2>&1
tail -3 gpurun_out/r02u_pytest.txt
run() {
  name=$1; shift
  timeout 400 python bench.py "$@" > gpurun_out/r02u_$name.json 2> gpurun_out/r02u_$name.err
  python - "$name" <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02u_{n}.json').read().strip().splitlines()[-1])
    print(n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'ingest',d['config'].get('frame_ingest'))
    ks=d['roofline']['kernels']
    print('  kernels us/launch:', {k: round(v['ms']*1000/v['calls'],1) for k,v in ks.items()})
    hp=d.get('host_phase_ms_per_frame_step') or {}
    nb=d['config']['batches_per_gpu']
    if 'prof' in n:
        print('  per batch-frame ms:', {k: round(v/nb,3) for k,v in sorted(hp.items()) if not k.startswith('x_')})
        print('  per sequence-frame us:', {k: round(v*1000/d['config']['sequences_per_gpu'],2) for k,v in sorted(hp.items()) if k.startswith('x_') and not k.startswith('x_i_')})
except Exception as e:
    print(n,'FAILED',e); print(open(f'gpurun_out/r02u_{n}.err').read()[-1200:])
P
}
run default --gpus 1 --steps 20 --warmup 5
run hostprof_e2e --steps 10 --warmup 3 --no-cpu-baseline --no-single-stream --profile-e2e --profile-overlapped --profile-level 3
run kprof_e2e --steps 10 --warmup 3 --no-cpu-baseline --no-single-stream --profile-e2e
