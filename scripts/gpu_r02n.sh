# round 2, GPU call 14: 1-point RANSAC on the device path against the reference's own estimator; whole GPU suite; bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_estimator.py -m gpu -q -rf --timeout 300 -p no:cacheprovider -k "matches_the_reference" > gpurun_out/r02n_pytest_ransac.txt 2>&1
tail -8 gpurun_out/r02n_pytest_ransac.txt
timeout 1500 python -m pytest tests -m gpu -q -rf --timeout 600 -p no:cacheprovider -x > gpurun_out/r02n_pytest_gpu.txt 2>&1
tail -5 gpurun_out/r02n_pytest_gpu.txt
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-single-stream > gpurun_out/r02n_default.json 2> gpurun_out/r02n_default.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02n_default.json').read().strip().splitlines()[-1])
print('default value',round(d['value']),'e2e',round(d['e2e']['value']),'ingest',d['config'].get('frame_ingest'),d['config'].get('frame_ingest_calibration_ms_per_step'))
P
