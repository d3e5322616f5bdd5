# usage: bash scripts/sweep_r01f.sh (GPU box) -- sequences x batches sweep of bench.py after the IMU kernel got 2x faster (diagnostic, not a bench line)
mkdir -p gpurun_out
one() { B=$1; NB=$2; shift 2
  env "$@" timeout 110 python bench.py --steps 20 --no-cpu-baseline --seqs $B --batches $NB 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B', $B, 'NB', d['config']['batches_per_gpu'], 'T', d['config']['host_threads'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), 'busy', round(d['roofline']['device_busy_frac'],2))" | tee -a gpurun_out/r01f_sweep.txt; }
one 256 4
one 256 8
one 384 6
one 512 8
one 512 4
