# round 2, GPU call 13: tensor-core downdate with the 512-thread streaming epilogue: parity, timing, ncu
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ekf.py -m gpu -q -rf --timeout 300 -p no:cacheprovider -k "tensor or batched" > gpurun_out/r02m_pytest_ekf.txt 2>&1
tail -4 gpurun_out/r02m_pytest_ekf.txt
timeout 600 python scripts/kbench_update.py 512 2>&1 | tee gpurun_out/r02m_kbench_update.txt
cat > /tmp/tc_one.py <<"P"
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from xivo_b200 import capi
ctx = capi.Context(0)
rng = np.random.default_rng(0)
N, M, B = 299, 124, 512
A = rng.normal(size=(N, N)); P1 = A @ A.T / N + np.eye(N)
P = np.broadcast_to(P1, (B, N, N)).copy()
H = rng.normal(size=(B, M, N)) * (rng.uniform(size=(B, M, N)) < 0.09)
ctx.ekf_update_batch(H, P, rng.normal(size=(B, M)), np.ones((B, M)), tf32x3=True, repeat=2)
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ekf_cov_tc2" -c 2 -o gpurun_out/r02m_tc2 -f python /tmp/tc_one.py > gpurun_out/r02m_ncu_tc2.log 2>&1
tail -2 gpurun_out/r02m_ncu_tc2.log
