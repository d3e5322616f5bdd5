"""Bring-up probe for the tcgen05 covariance downdate (GPU box): error of the tf32x3 update vs the fp64
kernels, normalised by sqrt(P_ii P_jj).  Usage: python scripts/tc_probe.py  (XIVO_TC_SWAP was a bring-up switch for the descriptor
roles; variant 0 is the one that is right and the only one left)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xivo_b200 import Context  # noqa: E402

ctx = Context(0)
worst = 0.0
for N, M in [(89, 24), (23, 2), (203, 60), (299, 124), (130, 33)]:
    rng = np.random.default_rng(N + M)
    A = rng.normal(size=(N, N))
    scale = np.exp(rng.uniform(-5, 1, N))
    P = (A @ A.T / N + np.eye(N)) * np.outer(scale, scale)
    P = 0.5 * (P + P.T)
    H = rng.normal(size=(M, N)) * (rng.uniform(size=(M, N)) < 0.15)
    inn = rng.normal(size=M)
    diagR = rng.uniform(0.5, 2.0, M)
    P64, e64 = ctx.ekf_update(H, P, inn, diagR)
    try:
        P32, e32 = ctx.ekf_update(H, P, inn, diagR, tf32x3=True)
    except Exception as ex:  # noqa: BLE001
        print("N", N, "M", M, "FAILED:", ex)
        worst = float("inf")
        break
    d = np.sqrt(np.outer(np.diag(P), np.diag(P)))
    rel = np.abs(P32 - P64) / d
    upd = np.abs(P64 - P) / d
    print("N %d M %d swap=%s: max |P32-P64|/sqrt(PiiPjj) = %.3e (update itself %.3e), symmetric=%s, err equal=%s" % (
        N, M, os.environ.get("XIVO_TC_SWAP", "0"), rel.max(), upd.max(), np.array_equal(P32, P32.T), np.array_equal(e32, e64)), flush=True)
    worst = max(worst, float(rel.max()) if np.isfinite(rel).all() else float("inf"))
ctx.close()
print("PROBE", "PASS" if worst <= 1e-5 else "FAIL", worst)
sys.exit(0 if worst <= 1e-5 else 3)
