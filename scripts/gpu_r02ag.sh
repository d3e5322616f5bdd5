# round 2: the from-scratch build of the final sources: smoke + workload test + a short bench line
mkdir -p gpurun_out
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 100 python -m pytest tests/test_gpu_bench_workload.py tests/test_gpu_tracker_decisions.py -x -q -m gpu 2>&1 | tail -1
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-single-stream > gpurun_out/r02ag_default.json 2> gpurun_out/r02ag_default.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02ag_default.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'launches',d['gpu_launches'])
P
