# usage: bash scripts/sweep_r01g.sh (GPU box) -- two larger points of the sequences x batches sweep, then the default bench line (512 x 8)
mkdir -p gpurun_out
one() { B=$1; NB=$2; shift 2
  env "$@" timeout 60 python bench.py --steps 20 --no-cpu-baseline --seqs $B --batches $NB 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B', $B, 'NB', d['config']['batches_per_gpu'], 'T', d['config']['host_threads'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), 'busy', round(d['roofline']['device_busy_frac'],2))" | tee -a gpurun_out/r01g_sweep.txt; }
one 768 8
one 1024 8
timeout 120 python bench.py > gpurun_out/r01g_bench.json 2> gpurun_out/r01g_bench.err
head -c 400 gpurun_out/r01g_bench.json
