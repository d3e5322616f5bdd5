# round 2, GPU call 12: timing of the update kernels at the BASELINE sizes x 512 filters; ncu of the two tensor-core downdate kernels at (299, 124)
mkdir -p gpurun_out
timeout 600 python scripts/kbench_update.py 512 2>&1 | tee gpurun_out/r02l_kbench_update.txt
cat > /tmp/tc_one.py <<'P'
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
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ekf_cov_tc2|ekf_gain" -c 4 -o gpurun_out/r02l_tc2 -f python /tmp/tc_one.py > gpurun_out/r02l_ncu_tc2.log 2>&1
tail -2 gpurun_out/r02l_ncu_tc2.log
