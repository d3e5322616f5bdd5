"""Generates tests/golden/ate_reference.npz by running the REFERENCE's own evaluation code
(/root/reference/scripts/tum_rgbd_benchmark_tools/{evaluate_ate,associate}.py, imported unmodified) on seeded
synthetic trajectories.  Only runs in the authoring container (the reference is not on the GPU box); the vectors
are what tests/test_dataio.py pins xivo_b200.dataio against."""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference/scripts/tum_rgbd_benchmark_tools")
if not hasattr(np.linalg, "linalg"):  # numpy >= 2: the reference spells numpy.linalg.linalg.svd
    import types

    np.linalg.linalg = types.SimpleNamespace(svd=np.linalg.svd)
import associate  # noqa: E402
import evaluate_ate  # noqa: E402

out = {}
for case in range(3):
    rng = np.random.default_rng(100 + case)
    n = 200 + 50 * case
    t_gt = np.sort(rng.uniform(0, 20, n))
    s = np.linspace(0, 4 * np.pi, n)
    gt = np.stack([np.cos(s) * (1 + 0.1 * case), np.sin(2 * s), 0.3 * s], 1) + rng.normal(0, 0.01, (n, 3))
    # estimate: rigidly moved, noisy, stamps jittered and subsampled
    keep = np.sort(rng.choice(n, n - 30, replace=False))
    A = rng.normal(size=(3, 3))
    Q, _ = np.linalg.qr(A)
    if np.linalg.det(Q) < 0:
        Q[:, 0] = -Q[:, 0]
    est = (gt[keep] @ Q.T + rng.normal(size=3)) + rng.normal(0, 0.05, (len(keep), 3))
    t_est = t_gt[keep] + rng.uniform(-0.015, 0.015, len(keep))
    first = {float(t): list(p) for t, p in zip(t_gt, gt)}
    second = {float(t): list(p) for t, p in zip(t_est, est)}
    matches = associate.associate(first, second, 0.0, 0.02)
    first_xyz = np.matrix([[float(v) for v in first[a][0:3]] for a, b in matches]).transpose()
    second_xyz = np.matrix([[float(v) for v in second[b][0:3]] for a, b in matches]).transpose()
    rot, trans, err = evaluate_ate.align(second_xyz, first_xyz)
    out[f"t_gt{case}"], out[f"gt{case}"], out[f"t_est{case}"], out[f"est{case}"] = t_gt, gt, t_est, est
    out[f"pairs{case}"] = np.array([[np.searchsorted(t_gt, a), int(np.nonzero(t_est == b)[0][0])] for a, b in matches])
    out[f"rot{case}"], out[f"trans{case}"] = np.asarray(rot), np.asarray(trans)
    out[f"rmse{case}"] = np.sqrt(np.dot(err, err) / len(err))
    out[f"stats{case}"] = np.array([np.mean(err), np.median(err), np.max(err)])
    print(case, len(matches), out[f"rmse{case}"])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ate_reference.npz"), **out)
