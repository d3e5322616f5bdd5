# round 2: is the CPU quota throttling the run?  cpu.stat around the timed passes, headroom sweep
mkdir -p gpurun_out
cat /sys/fs/cgroup/cpu.max > gpurun_out/r02ac_cpumax.txt; cat /sys/fs/cgroup/cpu.stat >> gpurun_out/r02ac_cpumax.txt
run() {
  name=$1; shift
  timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-single-stream "$@" > gpurun_out/r02ac_$name.json 2> gpurun_out/r02ac_$name.err
  python - "$name" <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02ac_{n}.json').read().strip().splitlines()[-1])
    print(n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'batches',d['config']['batches_per_gpu'],'threads',d['config'].get('host_threads'), d.get('host_cpu'))
except Exception as e:
    print(n,'FAILED',e); print(open(f'gpurun_out/r02ac_{n}.err').read()[-1500:])
P
}
run h1
run h3 --cpu-headroom 3
run h5 --cpu-headroom 5
XIVO_SPIN_MS=0 run spin0
