# round 2, GPU call 3: debug of the device-decision divergence + validation of the LK / TMA pyramid rewrite
mkdir -p gpurun_out
timeout 300 python scripts/debug_decisions.py > gpurun_out/r02c_debug.txt 2>&1; cat gpurun_out/r02c_debug.txt | tail -20
XIVO_HOST_TRACKER_DECISIONS=1 timeout 300 python -m pytest tests/test_gpu_estimator.py -m gpu -q -x -k "image_pipeline_parity" -p no:cacheprovider 2>&1 | tail -3
XIVO_PYRDOWN_TMA=0 timeout 300 python -m pytest tests/test_gpu_estimator.py -m gpu -q -x -k "image_pipeline_parity" -p no:cacheprovider 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q -rf --timeout 600 -p no:cacheprovider > gpurun_out/r02c_pytest_gpu.txt 2>&1
tail -12 gpurun_out/r02c_pytest_gpu.txt
run() {  # name, env..., -- bench args
  name=$1; shift
  envs=""; while [ "$1" != "--" ]; do envs="$envs $1"; shift; done; shift
  env $envs timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-single-stream "$@" > gpurun_out/r02c_$name.json 2> gpurun_out/r02c_$name.err
  python - "$name" <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02c_{n}.json').read().strip().splitlines()[-1])
    print(n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'lanes',d['config'].get('lanes_per_batch'),'ingest',d['config'].get('frame_ingest'),d['config'].get('frame_ingest_calibration_ms_per_step'),'tracked',round(d['tracked_features_mean']))
except Exception as e:
    print(n,'FAILED',e); print(open(f'gpurun_out/r02c_{n}.err').read()[-1500:])
P
}
run lanes_default --
run lanes_default_notma XIVO_PYRDOWN_TMA=0 --
run pool8 XIVO_LANES=1 -- --batches 8
run pool8_notma XIVO_LANES=1 XIVO_PYRDOWN_TMA=0 -- --batches 8
run pool8_hostdec XIVO_LANES=1 XIVO_HOST_TRACKER_DECISIONS=1 -- --batches 8
run pool4x128 XIVO_LANES=1 -- --batches 4
CMD="python bench.py --seqs 64 --batches 1 --steps 1 --warmup 3 --frames-per-step 2 --no-cpu-baseline --no-single-stream --ingest copy_engine"
XIVO_LANES=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"lk_kernel|pyrdown|fast_|track_|imu_cov|ekf_gain|ekf_cov|jacobian|subfilter|cov_edit|pack_state" -s 450 -c 30 -o gpurun_out/r02c_top -f $CMD > gpurun_out/r02c_top.log 2>&1
tail -2 gpurun_out/r02c_top.log
