# round 2, GPU call 4 (debug): TMA pyramid fault under compute-sanitizer; device-decision divergence with the TMA pass off
mkdir -p gpurun_out
timeout 120 python scripts/tma_probe.py > gpurun_out/r02d_probe_plain.txt 2>&1; tail -3 gpurun_out/r02d_probe_plain.txt
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python scripts/tma_probe.py > gpurun_out/r02d_sanitizer.txt 2>&1; grep -v "^=========     at\|Host Frame\|^=========         \|^$" gpurun_out/r02d_sanitizer.txt | head -40
XIVO_PYRDOWN_TMA=0 timeout 300 python scripts/debug_decisions.py > gpurun_out/r02d_debug.txt 2>&1; tail -22 gpurun_out/r02d_debug.txt
XIVO_PYRDOWN_TMA=0 XIVO_HOST_TRACKER_DECISIONS=1 timeout 300 python -m pytest tests/test_gpu_estimator.py -m gpu -q -x -k "image_pipeline_parity" -p no:cacheprovider 2>&1 | tail -3
XIVO_PYRDOWN_TMA=0 timeout 600 python -m pytest tests/test_gpu_tracker.py tests/test_gpu_tracker_decisions.py tests/test_gpu_ekf.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5
