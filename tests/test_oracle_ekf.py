"""Pins oracle/ekf_oracle.py with the reference's own unit-test methods (finite differences,
known answers) — src/test/unittest_jacobians_instate.cpp, unittest_camera_*.cpp,
unittest_givens.cpp.  CPU only."""
import math

import numpy as np
import pytest

from oracle import ekf_oracle as E
from xivo_b200 import synth


def fd_setup(seed=0):
    """unittest_jacobians_instate.cpp:22-74: perfect pinhole 640x480 f=580, pixel (25,46), log-depth 2."""
    rng = np.random.default_rng(seed)
    cam = E.Camera(0, 480, 640, 580.0, 580.0, 320.0, 240.0)
    lay = E.Layout(15, 30)
    Rsb, Tsb = synth.random_rotation(rng), rng.normal(0, 1, 3)
    Rbc, Tbc = synth.random_rotation(rng), rng.normal(0, 1, 3)
    Rr, Tr = synth.random_rotation(rng), rng.normal(0, 1, 3)
    xc = cam.unproject(np.array([25.0, 46.0]))
    x = np.array([xc[0], xc[1], 2.0])
    return cam, lay, Rsb, Tsb, Rbc, Tbc, Rr, Tr, x


def xcn_of(Rsb, Tsb, Rbc, Tbc, Rr, Tr, x):
    Xc, _ = E.unproject_logz(x)
    Xs = Rr @ (Rbc @ Xc + Tbc) + Tr
    return Rbc.T @ (Rsb.T @ (Xs - Tsb) - Tbc)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_instate_jacobian_finite_difference(seed):
    cam, lay, Rsb, Tsb, Rbc, Tbc, Rr, Tr, x = fd_setup(seed)
    _, _, c = E.feature_jacobian(lay, cam, Rsb, Tsb, Rbc, Tbc, Rr, Tr, x, np.zeros(2), 2, 3)
    d = 1e-6
    base = xcn_of(Rsb, Tsb, Rbc, Tbc, Rr, Tr, x)
    tol = 9e-4  # unittest_jacobians_instate.cpp tolerance
    for j in range(3):
        e = np.zeros(3)
        e[j] = d
        num = {
            "dXcn_dWsb": (xcn_of(Rsb @ E.so3_exp(e), Tsb, Rbc, Tbc, Rr, Tr, x) - base) / d,
            "dXcn_dTsb": (xcn_of(Rsb, Tsb + e, Rbc, Tbc, Rr, Tr, x) - base) / d,
            "dXcn_dWbc": (xcn_of(Rsb, Tsb, Rbc @ E.so3_exp(e), Tbc, Rr, Tr, x) - base) / d,
            "dXcn_dTbc": (xcn_of(Rsb, Tsb, Rbc, Tbc + e, Rr, Tr, x) - base) / d,
            "dXcn_dWsbr": (xcn_of(Rsb, Tsb, Rbc, Tbc, Rr @ E.so3_exp(e), Tr, x) - base) / d,
            "dXcn_dTsbr": (xcn_of(Rsb, Tsb, Rbc, Tbc, Rr, Tr + e, x) - base) / d,
            "dXcn_dx": (xcn_of(Rsb, Tsb, Rbc, Tbc, Rr, Tr, x + e) - base) / d,
        }
        for k, v in num.items():
            assert np.abs(c[k][:, j] - v).max() < tol * max(1.0, np.abs(v).max()), k


@pytest.mark.parametrize("model", [0, 3])
def test_camera_roundtrip_and_fd(model):
    """unittest_camera_pinhole.cpp / unittest_camera_equi.cpp: project∘unproject and FD Jacobian."""
    prob = synth.random_filter_problem(4, 8, 4, seed=0, model=model)
    cam = synth.camera_from_array(E, prob["camera"])
    rng = np.random.default_rng(0)
    for _ in range(20):
        xc = rng.uniform(-0.5, 0.5, 2)
        xp, J = cam.project(xc)
        assert np.abs(cam.unproject(xp) - xc).max() < 1e-8
        d = 1e-6
        for j in range(2):
            e = np.zeros(2)
            e[j] = d
            num = (cam.project(xc + e)[0] - xp) / d
            assert np.abs(num - J[:, j]).max() < 1e-3


def test_givens_known_answers():
    """unittest_givens.cpp:15-37 (values from the lecture note the reference cites)."""
    c, s = E.givens_pair(0.9134, 0.6324)
    assert abs(abs(c) - 0.8222) < 5e-4 and abs(abs(s) - 0.5692) < 5e-4
    c, s = E.givens_pair(0.1270, 1.1109)
    assert abs(abs(c) - 0.1136) < 5e-4 and abs(abs(s) - 0.9935) < 5e-4
    # G^T [a, b]^T = [r, 0]^T
    for a, b in ((0.9134, 0.6324), (0.1270, 1.1109), (-2.0, 0.3)):
        c, s = E.givens_pair(a, b)
        Gm = np.array([[c, s], [-s, c]])
        assert abs((Gm.T @ np.array([a, b]))[1]) < 1e-12


def test_joseph_equals_standard_form_and_psd():
    prob = synth.random_filter_problem(4, 14, 10, seed=1)
    synth.set_measurements_near_prediction(prob, E, 1.0)
    lay = E.Layout(4, 14)
    cam = synth.camera_from_array(E, prob["camera"])
    X = prob["X24"]
    H = np.zeros((20, lay.N))
    inn = np.zeros(20)
    for i in range(10):
        g = prob["groups"][prob["feat_ref"][i]]
        J, r, _ = E.feature_jacobian(lay, cam, X[:9].reshape(3, 3), X[9:12], X[12:21].reshape(3, 3), X[21:24], g[:9].reshape(3, 3),
                                     g[9:12], prob["feat_x"][i], prob["feat_xp"][i], int(prob["feat_ref"][i]), int(prob["feat_sind"][i]))
        E.fill_jacobian_block(lay, H, 2 * i, J, int(prob["feat_ref"][i]), int(prob["feat_sind"][i]))
        inn[2 * i : 2 * i + 2] = r
        # the FillJacobianBlock quirk: rotation block <- translation block, translation block zero
        go = lay.goff(int(prob["feat_ref"][i]))
        assert np.array_equal(H[2 * i : 2 * i + 2, go : go + 3], J[:, go + 3 : go + 6])
        assert np.all(H[2 * i : 2 * i + 2, go + 3 : go + 6] == 0)
    P = prob["P"]
    Pn, err, K, S = E.update_joseph(H, P, inn, np.ones(20))
    Pstd = P - K @ H @ P
    assert np.abs(Pn - Pstd).max() < 1e-10 * np.abs(P).max()
    assert np.abs(Pn - Pn.T).max() < 1e-12
    live = np.abs(P).sum(0) > 0
    w = np.linalg.eigvalsh(Pn[np.ix_(live, live)])
    assert w.min() > -1e-12
    assert np.allclose(err, K @ inn)


def test_slot_surgery_matches_reference_semantics():
    lay = E.Layout(4, 14)
    rng = np.random.default_rng(0)
    A = rng.normal(size=(lay.N, lay.N))
    P = A @ A.T
    err = np.zeros(lay.N)
    E.add_group_to_state(lay, P, err, 2)
    o = lay.goff(2)
    assert np.allclose(P[o : o + 3, o : o + 3], P[0:3, 0:3]) and np.allclose(P[o + 3 : o + 6, o + 3 : o + 6], P[3:6, 3:6])
    assert np.allclose(P[o : o + 3, 3:6], P[0:3, 3:6]) and np.allclose(P, P.T)
    E.add_feature_to_state(lay, P, 5, np.diag([1.0, 2.0, 3.0]))
    f = lay.foff(5)
    assert np.allclose(P[f : f + 3, f : f + 3], np.diag([1.0, 2.0, 3.0])) and np.abs(P[f : f + 3, :f]).max() == 0
    E.fix_feature_xy(lay, P, 5)
    assert P[f + 2, f + 2] == 3.0 and P[f, f] == 0
    E.remove_group_from_state(lay, P, err, 2)
    assert np.abs(P[o : o + 6]).max() == 0 and np.abs(P[:, o : o + 6]).max() == 0


def test_propagation_keeps_symmetry_and_psd():
    rng = np.random.default_rng(0)
    X = E.MotionState(np.eye(3), np.zeros(3), np.array([0.1, 0.0, 0.0]), np.zeros(3), np.zeros(3), np.eye(3), np.zeros(3), np.eye(3))
    Pmm = np.diag(rng.uniform(1e-4, 1e-2, 23))
    Qimu = np.diag([1e-4] * 3 + [1e-3] * 3 + [0] * 6)
    g = np.array([0, 0, -9.8])
    for method in ("PrinceDormand", "RK4"):
        Xc = X.copy()
        Phi, Pn = E.integrate(method, Xc, Pmm.copy(), np.array([0.01, 0.02, 0.03]), np.array([0.0, 0.0, 9.8]), np.zeros(3), np.zeros(3),
                              0.005, np.eye(3), np.eye(3), g, Qimu)
        assert np.abs(Pn - Pn.T).max() < 1e-15
        assert np.linalg.eigvalsh(Pn).min() > 0
        assert np.abs(Xc.Rsb @ Xc.Rsb.T - np.eye(3)).max() < 1e-12
        assert np.abs(Phi - np.eye(23)).max() < 0.1
        assert abs(Xc.Tsb[0] - 0.1 * 0.005) < 1e-6


# ------------------------------------------------------------------------------------------------------------------------
# The reference's own known-answer test of the angular triangulation methods (src/test/unittest_triangulation.cpp:18-206)
# ------------------------------------------------------------------------------------------------------------------------
def reference_triangulation_fixtures():
    """(name, xc1, z1, g21 (4x4), noise on xc2, expected return value) exactly as written in unittest_triangulation.cpp."""
    return [
        ("Normal_Inputs", (0.4, 0.6), 5.0, [[0.9849082, 0, 0.1731, -9.8490], [0, 1, 0, 0], [-0.17310, 0, 0.98490, 1.73101], [0, 0, 0, 1]], 0.0, True),
        ("Parallax", (2.2, 0.7), 5.0, [[0.9998, 0, 0.01745, -0.01], [0, 1, 0, 0], [-0.01745, 0, 0.9998, 0], [0, 0, 0, 1]], 0.0, False),
        ("Cheirality", (2.0, -0.77), 5.0, [[-1, 0, 0, 3], [0, 1, 0, 0], [0, 0, -1, 0], [0, 0, 0, 1]], 0.0, False),
        ("Angular_Reprojection_Error", (2.22216, 0.778023), 5.0, [[1, 0, 0, 3], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], 0.7, False),
        ("Vanishing_Point", (0.2, 0.3), 6000.0, [[1, 0, 0, -0.1], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], 0.0, False),
    ]


def reference_triangulation_inputs(xc1, z1, g21, noise):
    g21 = np.array(g21, float)
    X2 = g21 @ np.array([xc1[0] * z1, xc1[1] * z1, z1, 1.0])
    xc2 = np.array([X2[0] / X2[2] + np.float32(noise), X2[1] / X2[2] + np.float32(noise)])  # `float noise = 0.7`
    g12 = np.linalg.inv(g21)
    U, _, Vt = np.linalg.svd(g12[:3, :3])  # SE3::fitToSE3: nearest rotation
    return U @ Vt, g12[:3, 3].copy(), np.array(xc1, float), xc2


@pytest.mark.parametrize("name,xc1,z1,g21,noise,expect", reference_triangulation_fixtures(), ids=[f[0] for f in reference_triangulation_fixtures()])
def test_reference_triangulation_known_answers(name, xc1, z1, g21, noise, expect):
    """Thresholds of the fixture: 0.1 deg / 0.25 deg, depth within 0.5 of z1.  Angular_Reprojection_Error is the case the reference's own
    comment (unittest_triangulation.cpp:151-153) reports as failing for L1Angular in release builds: acos(1 + ulp) = NaN bypasses the check;
    with the clamped cosine (documented deviation) L1 is rejected as the test intends."""
    R12, t12, x1, x2 = reference_triangulation_inputs(xc1, z1, g21, noise)
    th, beta = np.float32(0.1 * math.pi / 180), np.float32(0.25 * math.pi / 180)
    for fn in (E.tri_l1_angular, E.tri_l2_angular, E.tri_linf_angular):
        ok, X = fn(R12, t12, x1, x2, th, beta)
        assert ok == expect, f"{name}: {fn.__name__}"
        if expect:
            assert abs(X[2] - z1) <= 0.5
