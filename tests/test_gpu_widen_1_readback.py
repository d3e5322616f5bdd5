"""Drop-in boundary on the GPU: every read-back method of the reference's Python binding (pybind11/pyxivo.cpp:332-398) through
xivo_b200.pyxivo.Estimator, at the end of a pinned point-cloud sequence, against what the REFERENCE'S OWN accessors returned for
that sequence (oracle/ref_wrap.cpp -> tests/golden/reference_pcw.npz `acc.*`, live where oracle/_ref is built).  The host-side
table logic is also checked on the CPU (tests/test_host_twin.py); here the covariance columns come from the device-resident P.
(File name sorts after the other GPU suites on purpose: this row was added after the last GPU minute of round 1.)"""
import numpy as np
import pytest

import test_reference_pin as RP
from xivo_b200 import pyxivo, sim

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["small_89", "default_203"])
def test_every_read_back_method_matches_the_reference(case, tmp_path):
    name, G, F, duration, seed, sim_depths, over, offset = next(c for c in RP.CASES if c[0] == case)
    cfg = sim.load_cfg(RP.CFG)
    ref, how = RP.reference_result(name, over, G, F, duration, seed, sim_depths, offset, tmp_path)
    msgs, _ = RP.stream(cfg, duration, seed, offset)
    e = pyxivo.Estimator(cfg, max_groups=G, max_features=F)
    e.InitWithSimDepths()
    for kind, ts, p in msgs:
        if kind == "imu":
            e.InertialMeas(ts, *p[0], *p[1])
        else:
            e.VisualMeasPointCloud(ts, p[0], p[1])
    a = lambda k: ref["acc." + k]
    # no-argument overloads: same rows; the reference's order is its raw-pointer order (changes between runs of the same binary)
    ids = e.InstateFeatureIDs().tolist()
    assert sorted(ids) == sorted(a("all.ids").tolist()) and len(ids) == e.num_instate_features()
    perm = [a("all.ids").tolist().index(i) for i in ids]
    cols = (("sinds", e.InstateFeatureSinds), ("refs", e.InstateFeatureRefGroups), ("Xs", e.InstateFeaturePositions), ("Xc", e.InstateFeatureXc),
            ("xc", e.InstateFeaturexc), ("pred", e.InstateFeaturePreds), ("meas", e.InstateFeatureMeas), ("cov", e.InstateFeatureCovs))
    for key, fn in cols:
        assert np.abs(fn() - a("all." + key)[perm]).max() <= 1e-7, f"{how}: {key}"
    # (int n_output) overloads: sorted by covariance norm, max(count, n) rows
    for tag, n in (("top5", 5), ("top50", 50)):
        k = len(a(tag + ".ids"))
        got = e.InstateFeatureIDs(n)
        assert len(got) == int(a(tag + ".rows")) and got[:k].tolist() == a(tag + ".ids").tolist(), f"{how}: {tag} order"
        for key, fn in cols:
            assert np.abs(fn(n)[:k] - a(f"{tag}.{key}")).max() <= 1e-7, f"{how}: {tag}.{key}"
    assert e.InstateGroupIDs().tolist() == a("groups.ids").tolist() and e.InstateGroupSinds().tolist() == a("groups.sinds").tolist()
    assert np.abs(e.InstateGroupPoses() - a("groups.pose")).max() <= 1e-7
    gc = e.InstateGroupCovs()
    assert gc.shape == (len(a("groups.ids")), 21) and np.abs(gc[:, :6] - a("groups.cov6")).max() <= 1e-7 * max(1.0, np.abs(a("groups.cov6")).max())
    blocks = e.InstateGroupCovBlocks()
    P = e.P()
    for i, s in enumerate(e.InstateGroupSinds()):
        o = 23 + 6 * int(s)
        assert np.array_equal(blocks[i], P[o:o + 6, o:o + 6])
    assert sorted(e.JustDroppedFeatureIDs().tolist()) == sorted(a("just_dropped").tolist())
    assert np.array_equal(e.Ca(), a("Ca")) and np.array_equal(e.Cg(), a("Cg")) and e.td() == float(a("td"))
    assert np.abs(e.Rg() - a("Rsg")).max() <= 1e-9 and np.abs(e.Pstate() - a("Pstate")).max() <= 1e-7 * np.abs(a("Pstate")).max()
    assert np.array_equal(e.CameraIntrinsics(), a("intrinsics")) and e.CameraDistortionType() == int(a("distortion_type"))
    tc = a("tracker_counters")
    assert (e.num_tracker_outlier_rejected(), e.num_oneptransac_rejected()) == (int(tc[0]), int(tc[3]))
    t_ids = [i for i, _ in e.tracked_features_no_descriptor()]
    assert t_ids == a("tracked.ids").tolist()
    assert np.abs(np.array([p for _, p in e.tracked_features_no_descriptor()]) - a("tracked.xy")).max() <= 1e-12
    assert len(e.tracked_features()) == len(t_ids) and e.Visualize() is None and e.UsingLoopClosure() is False
    v = e.Vsb().copy()
    e.ScaleInitVelocity(4.0)
    assert np.array_equal(e.Vsb(), v / 4.0)
    e.close()


def test_cpp_facade_read_back_surface(tmp_path):
    """include/xivo_b200.hpp: every new accessor of the C++ facade executed on the GPU (tests/cpp/facade_check.cpp --readback)."""
    import subprocess

    import test_cpp_facade as TF

    exe = TF.build(tmp_path)
    r = subprocess.run([exe, RP.CFG, "--readback"], capture_output=True, text=True)
    assert r.returncode == 0 and "readback ok" in r.stdout, r.stdout + r.stderr
