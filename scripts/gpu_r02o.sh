# round 2, GPU call: descriptor path (BRIEF, rescue, MATCH tracker) against the oracle; then the whole GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_widen_5_descriptors.py -m gpu -q -rf --timeout 300 -p no:cacheprovider > gpurun_out/r02o_pytest_desc.txt 2>&1
tail -25 gpurun_out/r02o_pytest_desc.txt | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -q -rf --timeout 600 -p no:cacheprovider -x --deselect tests/test_gpu_widen_5_descriptors.py > gpurun_out/r02o_pytest_gpu.txt 2>&1
tail -5 gpurun_out/r02o_pytest_gpu.txt
