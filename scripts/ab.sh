# usage: bash scripts/ab.sh  -- A/B of host options on a GPU box (diagnostic)
one() { env "$@" python bench.py --steps 20 --no-cpu-baseline --seqs ${B:-192} --batches ${NB:-3} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', 'B', ${B:-192}, 'NB', ${NB:-3}, round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), 'busy', round(d['roofline']['device_busy_frac'],2))"; }
NB=4 B=256 one XIVO_ZEROCOPY=1
NB=4 B=256 one XIVO_ZEROCOPY=0
NB=4 B=256 one XIVO_ZEROCOPY=1
NB=3 B=192 one XIVO_ZEROCOPY=1
NB=4 B=512 one XIVO_ZEROCOPY=1
