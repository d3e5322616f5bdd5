# usage: [NB=n] bash scripts/prof_run.sh B [level]  -- one profiled bench run with host-phase breakdown (diagnostic)
B=${1:-128}; LV=${2:-1}
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 3 --seqs $B --batches ${NB:-1} --no-cpu-baseline --profile-level $LV $EXTRA > gpurun_out/s.json 2> gpurun_out/s.err
python - <<PY
import json
d=json.loads(open("gpurun_out/s.json").read().strip().splitlines()[-1])
B=$B
print("value %.0f e2e %.0f ms/step %.3f e2e ms/step %.3f profiled %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["profiled_pass_ms_per_step"]))
print("kern us/launch", {k:round(v["ms"]/v["calls"]*1000,1) for k,v in d["roofline"]["kernels"].items()})
hp=d["host_phase_ms_per_step"]
bl={k:v for k,v in hp.items() if not k.startswith("x_")}
print("batch-level ms/step:", bl)
print("per-seq us:", {k:round(v*1000/B,1) for k,v in hp.items() if k.startswith("x_")})
PY
