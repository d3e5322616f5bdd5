# round 2, GPU call 11: second tensor-core downdate kernel (TMA-staged TF32 operands): parity, then timing at the BASELINE sizes x 512 filters
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ekf.py -m gpu -q -rf --timeout 300 -p no:cacheprovider > gpurun_out/r02k_pytest_ekf.txt 2>&1
tail -15 gpurun_out/r02k_pytest_ekf.txt
timeout 600 python -m pytest tests/test_gpu_estimator.py -m gpu -q -rf --timeout 300 -p no:cacheprovider -k "tensor" > gpurun_out/r02k_pytest_tc_pipeline.txt 2>&1
tail -5 gpurun_out/r02k_pytest_tc_pipeline.txt
timeout 600 python scripts/kbench_update.py 512 2>&1 | tee gpurun_out/r02k_kbench_update.txt
