# round 2, GPU call 5: TMA variant probe, full suite (TMA pyramid off by default), bench sweep, ncu of every kernel
mkdir -p gpurun_out
for v in 0 1 2 3 4 5 6 7 8 9 10 11 12 13 14; do timeout 60 scripts/probes/tma_probe $v; done 2>&1 | tee gpurun_out/r02e_tma_probe.txt
timeout 1500 python -m pytest tests -m gpu -q -rf --timeout 600 -p no:cacheprovider --deselect tests/test_gpu_tracker_decisions.py::test_tma_pyramid_equals_thread_staged_pyramid > gpurun_out/r02e_pytest_gpu.txt 2>&1
tail -8 gpurun_out/r02e_pytest_gpu.txt
run() {  # name, env..., -- bench args
  name=$1; shift
  envs=""; while [ "$1" != "--" ]; do envs="$envs $1"; shift; done; shift
  env $envs timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-single-stream "$@" > gpurun_out/r02e_$name.json 2> gpurun_out/r02e_$name.err
  python - "$name" <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02e_{n}.json').read().strip().splitlines()[-1])
    print(n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'lanes',d['config'].get('lanes_per_batch'),'ingest',d['config'].get('frame_ingest'),d['config'].get('frame_ingest_calibration_ms_per_step'),'tracked',round(d['tracked_features_mean']))
except Exception as e:
    print(n,'FAILED',e); print(open(f'gpurun_out/r02e_{n}.err').read()[-800:])
P
}
run lanes_default --
run lanes_default_nozc XIVO_ZC_TABLES=0 --
run lanes16 XIVO_LANES=16 --
run pool8 XIVO_LANES=1 -- --batches 8
run pool8_nozc XIVO_LANES=1 XIVO_ZC_TABLES=0 -- --batches 8
run pool8_hostdec XIVO_LANES=1 XIVO_HOST_TRACKER_DECISIONS=1 -- --batches 8
run pool4x128 XIVO_LANES=1 -- --batches 4
run pool16x32 XIVO_LANES=1 -- --batches 16
CMD="python bench.py --seqs 64 --batches 1 --steps 1 --warmup 3 --frames-per-step 2 --no-cpu-baseline --no-single-stream --ingest copy_engine"
XIVO_LANES=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"lk_kernel|pyrdown|fast_|track_|imu_cov|ekf_gain|ekf_cov|jacobian|subfilter|cov_edit|pack_state" -s 450 -c 30 -o gpurun_out/r02e_top -f $CMD > gpurun_out/r02e_top.log 2>&1
tail -2 gpurun_out/r02e_top.log
