"""Python-side logic of xivo_b200.pyxivo.Estimator that needs no GPU: the (int n_output) overloads' row padding, the column mapping of
the per-feature table and the reference's InstateGroupCovs layout, on a stand-in batch fed from the reference's own accessor dump
(tests/golden/reference_pcw.npz `acc.*`).  The same methods run against the device in tests/test_gpu_widen_1_readback.py."""
import os

import numpy as np

from xivo_b200 import pyxivo

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_pcw.npz"))
A = {k[len("small_89.acc."):]: GOLD[k] for k in GOLD.files if k.startswith("small_89.acc.")}
KEYS = {"ids": "ids", "sinds": "sinds", "ref_groups": "refs", "Xs": "Xs", "Xc": "Xc", "xc": "xc", "pred": "pred", "meas": "meas", "cov": "cov"}


class FakeBatch:
    """What Batch.instate_feature_table / instate_group_table return, taken from the dump (top50 holds every in-state feature, sorted)."""

    def instate_feature_table(self, seq=0, n_output=-1):
        tag, k = ("all", None) if n_output < 0 else ("top50", min(n_output, len(A["top50.ids"])))
        return {ours: A[f"{tag}.{ref}"][:k].copy() for ours, ref in KEYS.items()}

    def instate_group_table(self, seq=0):
        n = len(A["groups.ids"])
        cov = np.arange(36.0 * n).reshape(n, 6, 6)
        return dict(ids=A["groups.ids"], sinds=A["groups.sinds"], pose=A["groups.pose"], cov=cov + cov.transpose(0, 2, 1))


def facade():
    e = pyxivo.Estimator.__new__(pyxivo.Estimator)
    e._b = FakeBatch()
    return e


def test_n_output_overloads_have_the_reference_row_count_and_order():
    e, count = facade(), len(A["top50.ids"])
    for n in (0, 5, count, 50):
        ids = e.InstateFeatureIDs(n)
        assert len(ids) == max(count, n) == (int(A["top50.rows"]) if n == 50 else max(count, n))  # estimator_accessors.cpp: npts = max(size, n_output)
        k = min(count, n)
        assert ids[:k].tolist() == A["top50.ids"][:k].tolist() and not ids[k:].any()  # rows past min(count, n) are never written
        assert e.InstateFeaturePositions(n).shape == (max(count, n), 3) and e.InstateFeatureCovs(n).shape == (max(count, n), 6)
        assert e.InstateFeaturePreds(n).shape == (max(count, n), 2)
    assert e.InstateFeatureIDs(5).tolist()[:5] == A["top5.ids"].tolist() and np.array_equal(e.InstateFeatureCovs(5)[:5], A["top5.cov"])


def test_no_argument_overloads_and_column_mapping():
    e = facade()
    for fn, key in ((e.InstateFeatureIDs, "ids"), (e.InstateFeatureSinds, "sinds"), (e.InstateFeatureRefGroups, "refs"), (e.InstateFeaturePositions, "Xs"),
                    (e.InstateFeatureXc, "Xc"), (e.InstateFeaturexc, "xc"), (e.InstateFeaturePreds, "pred"), (e.InstateFeatureMeas, "meas"), (e.InstateFeatureCovs, "cov")):
        assert np.array_equal(fn(), A["all." + key])
    assert np.array_equal(e.InstateGroupPoses(), A["groups.pose"]) and e.InstateGroupPoses().shape[1] == 7


def test_group_covs_reproduce_the_reference_column_bug():
    e = facade()
    blocks, out = e.InstateGroupCovBlocks(), e.InstateGroupCovs()
    assert out.shape == (len(blocks), 21) and not out[:, 6:].any()
    for i, c in enumerate(blocks):  # estimator_accessors.cpp InstateGroupCovs: `cnt = 0` inside the row loop
        assert out[i, :6].tolist() == [c[5, 5], c[4, 5], c[3, 5], c[2, 5], c[1, 5], c[0, 5]]
