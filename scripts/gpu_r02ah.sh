# round 2: BASELINE configs[3] (800 features: 808-track accept / select launches) with the final kernels
mkdir -p gpurun_out
timeout 105 python bench.py --config 3 --steps 2 --warmup 3 --no-cpu-baseline --no-single-stream --profile-steps 1 > gpurun_out/r02ah_config3.json 2> gpurun_out/r02ah_config3.err
echo rc=$?
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/r02ah_config3.json').read().strip().splitlines()[-1])
    print('value',round(d['value']),'e2e',round(d['e2e']['value']),'tracked',round(d['tracked_features_mean']))
    print({k: round(v['ms']*1e3/max(v['calls'],1),1) for k,v in d['roofline']['kernels'].items()})
except Exception as e:
    print('FAILED',e); print(open('gpurun_out/r02ah_config3.err').read()[-800:])
P
