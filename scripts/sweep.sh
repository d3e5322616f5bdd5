# usage: bash scripts/sweep.sh  -- host-configuration sweep of bench.py on a GPU box (diagnostic, not a bench line)
mkdir -p gpurun_out
thr() { grep nr_throttled /sys/fs/cgroup/cpu.stat | cut -d' ' -f2; }
run() {
  desc="$1"; shift
  t0=$(thr)
  env "$@" timeout 200 python bench.py --steps 20 --warmup 3 --seqs $B --batches $NB --no-cpu-baseline > gpurun_out/s.json 2> gpurun_out/s.err
  t1=$(thr)
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s.json").read().strip().splitlines()[-1])
    print("$desc B $B NB $NB T %s value %.0f e2e %.0f ms/step %.2f busy %.2f throttled %d" % (d["config"]["host_threads"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["device_busy_frac"], $t1-$t0))
except Exception as e:
    print("$desc failed", e); print(open("gpurun_out/s.err").read()[-600:])
PY
}
B=128; NB=1; run "pin2"
B=128; NB=1; run "pin2"
B=128; NB=1; run "pin1" XIVO_PIN=1
B=128; NB=1; run "pin0" XIVO_PIN=0
B=192; NB=3; run "pin2"
B=192; NB=3; run "pin2"
B=192; NB=3; run "pin1" XIVO_PIN=1
B=192; NB=3; run "pin0" XIVO_PIN=0
B=128; NB=2; run "pin2"
B=256; NB=4; run "pin2"
