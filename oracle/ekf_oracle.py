"""oracle/ekf_oracle.py — TEST INFRASTRUCTURE ONLY.

numpy fp64 restatement of the EKF side of XIVO's per-frame hot path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (xivo_b200/) never does.

Every function cites the reference lines it follows (paths relative to
/root/reference).  Pinning: the reference's unit tests for this path are
finite-difference identities (src/test/unittest_jacobians_instate.cpp:22-74,170-380,
unittest_jacobians_oos.cpp, unittest_camera_{pinhole,equi}.cpp) and the Givens
known answers (unittest_givens.cpp:15-37); tests/test_oracle_ekf.py re-runs those
on this restatement.  UpdateJosephForm / MHGating / SubfilterUpdate have no test in
the reference: for them parity is UNPINNED beyond algebraic
identities (Joseph == standard form for the optimal gain, symmetry, PSD).
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

# ----------------------------------------------------------------------------
# error-state layout, src/core.h:40-105 (default build flags: no online calib)
# ----------------------------------------------------------------------------
WSB, TSB, VSB, BG, BA, WBC, TBC, WSG = 0, 3, 6, 9, 12, 15, 18, 21
K_MOTION = 23


@dataclasses.dataclass(frozen=True)
class Layout:
    """kMaxGroup/kMaxFeature are compile-time in the reference (core.h:92-105)."""

    G: int = 15
    F: int = 30

    @property
    def group_begin(self) -> int:
        return K_MOTION

    @property
    def feature_begin(self) -> int:
        return K_MOTION + 6 * self.G

    @property
    def N(self) -> int:
        return K_MOTION + 6 * self.G + 3 * self.F

    def goff(self, sind: int) -> int:
        return self.group_begin + 6 * sind

    def foff(self, sind: int) -> int:
        return self.feature_begin + 3 * sind


# ----------------------------------------------------------------------------
# SO(3) helpers (Sophus SO3::exp / hat; helpers.cpp:374-378 SO3_from_rotvec)
# ----------------------------------------------------------------------------
def hat(w):
    return np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])


def so3_exp(w):
    """Rodrigues; matches Sophus SO3::exp to rounding (it goes through a quaternion)."""
    w = np.asarray(w, dtype=np.float64)
    th2 = float(w @ w)
    th = math.sqrt(th2)
    W = hat(w)
    if th < 1e-10:
        return np.eye(3) + W + 0.5 * W @ W
    return np.eye(3) + (math.sin(th) / th) * W + ((1.0 - math.cos(th)) / th2) * (W @ W)


def so3_log(R):
    c = max(-1.0, min(1.0, 0.5 * (np.trace(R) - 1.0)))
    th = math.acos(c)
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    if th < 1e-10:
        return 0.5 * v
    return th / (2.0 * math.sin(th)) * v


def so3_from_rotvec(w):
    """helpers.cpp:374-378: AngleAxis(|w|, w/|w|) -> quaternion -> SO3.
    Eigen's normalized() of the zero vector returns it unchanged -> identity."""
    return so3_exp(w)


# ----------------------------------------------------------------------------
# projections, common/project.h
# ----------------------------------------------------------------------------
def project(Xc):
    """common/project.h:11-24 -> (xc, dxc_dXc 2x3)."""
    X, Y, Z = Xc
    xc = np.array([X / Z, Y / Z])
    J = np.array([[1 / Z, 0, -X / (Z * Z)], [0, 1 / Z, -Y / (Z * Z)]])
    return xc, J


def unproject_logz(x):
    """common/project.h:79-95 -> (Xc, dXc_dx 3x3)."""
    with np.errstate(all="ignore"):
        z = float(np.exp(np.float64(x[2])))  # IEEE like the C library: exp of a diverged log-depth is inf, not an exception
    Xc = np.array([x[0] * z, x[1] * z, z])
    J = np.array([[z, 0, x[0] * z], [0, z, x[1] * z], [0, 0, z]])
    return Xc, J


def project_logz(Xc):
    """common/project.h:61-74."""
    X, Y, Z = Xc
    x = np.array([X / Z, Y / Z, math.log(Z)])
    J = np.array([[1 / Z, 0, -X / (Z * Z)], [0, 1 / Z, -Y / (Z * Z)], [0, 0, 1 / Z]])
    return x, J


# ----------------------------------------------------------------------------
# camera models
# ----------------------------------------------------------------------------
@dataclasses.dataclass
class Camera:
    """model 0 = pinhole (common/camera_pinhole.h:17-52),
    model 3 = equidistant (common/camera_equidist.h:23-160); enum values follow
    DistortionType (common/camera_base.h:13-18)."""

    model: int
    rows: int
    cols: int
    fx: float
    fy: float
    cx: float
    cy: float
    k: tuple = (0.0, 0.0, 0.0, 0.0)
    max_iter: int = 15

    def as_array(self):
        return np.array([self.model, self.rows, self.cols, self.fx, self.fy, self.cx, self.cy, *self.k], dtype=np.float64)

    def project(self, xc):
        if self.model == 0:
            xp = np.array([self.fx * xc[0] + self.cx, self.fy * xc[1] + self.cy])
            return xp, np.array([[self.fx, 0.0], [0.0, self.fy]])
        k0, k1, k2, k3 = self.k
        n2 = xc[0] * xc[0] + xc[1] * xc[1]
        n = math.sqrt(n2)
        n3 = n2 + 1
        th = math.atan2(n, 1.0)
        phi = math.atan2(xc[1], xc[0])
        th2 = th * th
        th3 = th2 * th
        th4 = th3 * th
        th5 = th3 * th2
        th6 = th5 * th
        th7 = th5 * th2
        th8 = th7 * th
        th9 = th7 * th2
        r = th + k0 * th3 + k1 * th5 + k2 * th7 + k3 * th9
        c, s = math.cos(phi), math.sin(phi)
        xp = np.array([self.fx * r * c + self.cx, self.fy * r * s + self.cy])
        dphi_dx, dphi_dy = -xc[1] / n2, xc[0] / n2
        dth_dx, dth_dy = xc[0] / n3 / n, xc[1] / n3 / n
        dr = 1 + k0 * 3 * th2 + k1 * 5 * th4 + k2 * 7 * th6 + k3 * 9 * th8
        J = np.array(
            [
                [self.fx * c * dr * dth_dx - self.fx * r * s * dphi_dx, self.fx * c * dr * dth_dy - self.fx * r * s * dphi_dy],
                [self.fy * s * dr * dth_dx + self.fy * r * c * dphi_dx, self.fy * s * dr * dth_dy + self.fy * r * c * dphi_dy],
            ]
        )
        return xp, J

    def unproject(self, xp):
        if self.model == 0:
            return np.array([(xp[0] - self.cx) / self.fx, (xp[1] - self.cy) / self.fy])
        k0, k1, k2, k3 = self.k
        xn, yn = xp[0] - self.cx, xp[1] - self.cy
        b, a = self.fx * yn, self.fy * xn
        phi = math.atan2(b, a)
        c, s = math.cos(phi), math.sin(phi)
        rth = xn / (self.fx * c)
        th = rth
        for _ in range(self.max_iter):
            th2 = th * th
            th3 = th2 * th
            th4 = th2 * th2
            th6 = th4 * th2
            x0 = k0 * th3 + k1 * th4 * th + k2 * th6 * th + k3 * th6 * th3 - rth + th
            x1 = 3 * k0 * th2 + 5 * k1 * th4 + 7 * k2 * th6 + 9 * k3 * th6 * th2 + 1
            d = 2 * x0 * x1
            d2 = 4 * th * x0 * (3 * k0 + 10 * k1 * th2 + 21 * k2 * th4 + 36 * k3 * th6) + 2 * x1 * x1
            th -= d / d2
        t = math.tan(th)
        return np.array([t * c, t * s])

    def focal_length(self):
        """camera_manager.cpp:56: fl_ = 0.5*sqrt(fx^2+fy^2) at construction."""
        return 0.5 * math.sqrt(self.fx * self.fx + self.fy * self.fy)


# ----------------------------------------------------------------------------
# in-state measurement Jacobian
# ----------------------------------------------------------------------------
def feature_jacobian(lay: Layout, cam: Camera, Rsb, Tsb, Rbc, Tbc, Rsbr, Tsbr, x, xp_meas, ref_sind, f_sind):
    """Feature::ComputeJacobian, src/feature.cpp:542-656 (USE_ONLINE_* off).
    Returns J (2 x N), inn (2), and the cache blocks the reference unit tests check."""
    Rsb_t, Rbc_t = Rsb.T, Rbc.T
    Xc, dXc_dx = unproject_logz(x)
    Xbr = Rbc @ Xc + Tbc
    Xs = Rsbr @ Xbr + Tsbr
    Xb = Rsb_t @ (Xs - Tsb)
    Xcn = Rbc_t @ (Xb - Tbc)
    dXbr_dXc = Rbc
    dXbr_dTbc = np.eye(3)
    dXbr_dWbc = -Rbc @ hat(Xc)
    dXs_dXbr = Rsbr
    dXs_dTsbr = np.eye(3)
    dXs_dWsbr = -Rsbr @ hat(Xbr)
    dXb_dXs = Rsb_t
    dXb_dTsb = -Rsb_t
    dXb_dWsb = hat(Xb)
    dXcn_dXb = Rbc_t
    dXcn_dTbc = -Rbc_t + dXcn_dXb @ dXb_dXs @ dXs_dXbr @ dXbr_dTbc
    dXcn_dWbc = hat(Xcn) + dXcn_dXb @ dXb_dXs @ dXs_dXbr @ dXbr_dWbc
    dXcn_dTsb = dXcn_dXb @ dXb_dTsb
    dXcn_dWsb = dXcn_dXb @ dXb_dWsb
    dXcn_dTsbr = dXcn_dXb @ dXb_dXs @ dXs_dTsbr
    dXcn_dWsbr = dXcn_dXb @ dXb_dXs @ dXs_dWsbr
    dXcn_dXs = dXcn_dXb @ dXb_dXs
    dXcn_dx = dXcn_dXs @ dXs_dXbr @ dXbr_dXc @ dXc_dx
    xcn, dxcn_dXcn = project(Xcn)
    xp, dxp_dxcn = cam.project(xcn)
    dxp_dXcn = dxp_dxcn @ dxcn_dXcn
    J = np.zeros((2, lay.N))
    J[:, WSB : WSB + 3] = dxp_dXcn @ dXcn_dWsb
    J[:, TSB : TSB + 3] = dxp_dXcn @ dXcn_dTsb
    J[:, WBC : WBC + 3] = dxp_dXcn @ dXcn_dWbc
    J[:, TBC : TBC + 3] = dxp_dXcn @ dXcn_dTbc
    goff, foff = lay.goff(ref_sind), lay.foff(f_sind)
    J[:, goff : goff + 3] = dxp_dXcn @ dXcn_dWsbr
    J[:, goff + 3 : goff + 6] = dxp_dXcn @ dXcn_dTsbr
    J[:, foff : foff + 3] = dxp_dXcn @ dXcn_dx
    inn = np.asarray(xp_meas, dtype=np.float64) - xp
    cache = dict(
        Xc=Xc, Xs=Xs, Xcn=Xcn, xcn=xcn, xp=xp, dXcn_dWsb=dXcn_dWsb, dXcn_dTsb=dXcn_dTsb, dXcn_dWbc=dXcn_dWbc,
        dXcn_dTbc=dXcn_dTbc, dXcn_dWsbr=dXcn_dWsbr, dXcn_dTsbr=dXcn_dTsbr, dXcn_dx=dXcn_dx, dxp_dXcn=dxp_dXcn,
    )
    return J, inn, cache


def fill_jacobian_block(lay: Layout, H, row, J, ref_sind, f_sind):
    """Feature::FillJacobianBlock, src/feature.cpp:658-684 — INCLUDING the
    reference's quirk at :675-676: the reference-group rotation block of H is
    overwritten with the translation block of J and H(:, goff+3:goff+6) stays 0."""
    for off in (WSB, TSB, WBC, TBC):
        H[row : row + 2, off : off + 3] = J[:, off : off + 3]
    goff, foff = lay.goff(ref_sind), lay.foff(f_sind)
    H[row : row + 2, goff : goff + 3] = J[:, goff : goff + 3]
    H[row : row + 2, goff : goff + 3] = J[:, goff + 3 : goff + 6]
    H[row : row + 2, foff : foff + 3] = J[:, foff : foff + 3]


def mh_distance(J, P, inn, R):
    """Estimator::MHGating distance, src/update.cpp:60-69: S = J P J^T + R I2, d = r^T S^-1 r."""
    S = J @ P @ J.T
    S[0, 0] += R
    S[1, 1] += R
    return float(inn @ np.linalg.solve(S, inn))


def mh_gating_select(dist, is_gauge, mh_thresh, mh_mult, min_inliers):
    """Threshold-relaxation loop of Estimator::MHGating, src/update.cpp:71-96.
    Returns (inlier_mask, num_rejected_accumulated).  NB the reference accumulates
    num_mh_rejected_ over relaxation rounds (update.cpp:88) — reproduced."""
    dist = np.asarray(dist)
    n = len(dist)
    thresh = mh_thresh
    num_rejected = 0
    inl = np.zeros(n, bool)
    while inl.sum() < min_inliers:
        inl = dist < thresh
        num_rejected += int((~inl).sum())
        thresh *= mh_mult
    return inl, num_rejected


def update_joseph(H, P, inn, diagR):
    """Estimator::UpdateJosephForm, src/estimator.cpp:1257-1288, same expression
    sequence: S=HPH^T+R; K^T = S^-1 (H P) (reference: LDLT); err = K inn;
    A = K H - I; P = A P A^T + (K sqrt(R))(K sqrt(R))^T."""
    HP = H @ P
    S = HP @ H.T + np.diag(diagR)
    Kt = np.linalg.solve(S, HP)
    K = Kt.T
    err = K @ inn
    A = K @ H - np.eye(P.shape[0])
    Pn = A @ P @ A.T
    Ks = K * np.sqrt(diagR)[None, :]
    Pn = Pn + Ks @ Ks.T
    return Pn, err, K, S


def subfilter_update(cam: Camera, x, P, xp_meas, gsb, gbc, gref, Rtri, mh_thresh, outlier_counter):
    """Feature::SubfilterUpdate, src/feature.cpp:246-297.  g* are (R, T) tuples.
    Returns (x, P, outlier_counter)."""
    Xc, dXc_dx = unproject_logz(x)
    Rsb, Tsb = gsb
    Rbc, Tbc = gbc
    Rr, Tr = gref
    # gtot = (gsb*gbc)^-1 * gref * gbc
    Rsc, Tsc = Rsb @ Rbc, Rsb @ Tbc + Tsb
    Rrc, Trc = Rr @ Rbc, Rr @ Tbc + Tr
    Rtot = Rsc.T @ Rrc
    Ttot = Rsc.T @ (Trc - Tsc)
    Xcn = Rtot @ Xc + Ttot
    xcn, dxcn_dXcn = project(Xcn)
    xp, dxp_dxcn = cam.project(xcn)
    H = dxp_dxcn @ dxcn_dXcn @ Rtot @ dXc_dx
    inn = np.asarray(xp_meas, dtype=np.float64) - xp
    S = H @ P @ H.T
    S[0, 0] += Rtri
    S[1, 1] += Rtri
    ratio = float(inn @ np.linalg.solve(S, inn)) / mh_thresh
    if ratio > 1:
        S[0, 0] += Rtri * (ratio - 1)
        S[1, 1] += Rtri * (ratio - 1)
        outlier_counter = outlier_counter + math.sqrt(ratio)
    else:
        outlier_counter = 0.0
    K = P @ H.T @ np.linalg.inv(S)
    xn = np.asarray(x, dtype=np.float64) + K @ inn
    I_KH = np.eye(3) - K @ H
    Pn = I_KH @ P @ I_KH.T + K @ (Rtri * K.T)
    return xn, Pn, outlier_counter


# -------------------------------------------------------------------------------------------------
# Two-view depth triangulation (Feature::Triangulate, src/feature.cpp:686-751; helpers.cpp:103-372).
# The reference stores several intermediates in `float` (a0/a1, the lambdas, the angles, the thresholds
# arrive as float parameters): those narrowings are reproduced with numpy.float32.
# -------------------------------------------------------------------------------------------------
_f32 = np.float32


def _acos_f32(c):
    """float theta = acos(double).  DOCUMENTED DEVIATION: the cosine is clamped to [-1, 1].  In L1Angular one of the two
    "reprojected" rays is the measured ray itself, so the reference evaluates acos(m.m / (|m| |m|)) — 1 up to rounding.  Whether
    that is 1 + ulp (acos = NaN, and a NaN theta0 makes `max_theta > thresh` false, i.e. the check is bypassed) or <= 1 is decided
    by the compiler's FMA contraction of Eigen's dot products (seen in the disassembly of the reference built here); the intended
    value is 0, which the clamp gives deterministically.  With a threshold the check cannot fail at (90 deg) the reference is
    reproduced to 1e-14 (tests/test_reference_pin.py)."""
    return _f32(math.acos(min(1.0, max(-1.0, c)))) if c == c else _f32(np.nan)


def _std_max(a, b):
    """std::max(a, b) = (a < b) ? b : a  (helpers.cpp:353)."""
    return b if a < b else a


def tri_checks(z, t10, m0, Rf0p, m1, f1p, max_theta, beta_thresh):
    """check_cheirality && check_angular_reprojection && check_parallax (helpers.cpp:330-372), short-circuit order kept."""
    zn2 = np.linalg.norm(z) ** 2
    with np.errstate(all="ignore"):
        lam0 = _f32(z @ np.cross(t10, f1p) / zn2)
        lam1 = _f32(z @ np.cross(t10, Rf0p) / zn2)
        if lam0 <= 0 or lam1 <= 0:
            return False
        th0 = _acos_f32(m0 @ Rf0p / (np.linalg.norm(m0) * np.linalg.norm(Rf0p)))
        th1 = _acos_f32(m1 @ f1p / (np.linalg.norm(m1) * np.linalg.norm(f1p)))
        if _std_max(th0, th1) > _f32(max_theta):
            return False
        beta = _acos_f32(f1p @ Rf0p / (np.linalg.norm(f1p) * np.linalg.norm(Rf0p)))
        if beta < _f32(beta_thresh):
            return False
    return True


def _angular_finish(R01, t01, t10, m0, m1, m0p, m1p, max_theta, beta_thresh):
    z = np.cross(m1p, m0p)
    with np.errstate(all="ignore"):
        X = (z @ np.cross(t10, m0p)) / np.linalg.norm(z) ** 2 * m1p
    X = R01 @ X + t01
    return tri_checks(z, t10, m0, m0p, m1, m1p, max_theta, beta_thresh), X


def _bearings(R01, t01, xc0, xc1):
    R10 = R01.T
    t10 = -1 * R01.T @ t01
    f0 = np.array([xc0[0], xc0[1], 1.0])
    f0 = f0 / np.linalg.norm(f0)
    f1 = np.array([xc1[0], xc1[1], 1.0])
    f1 = f1 / np.linalg.norm(f1)
    return t10, R10 @ f0, f1


def tri_l1_angular(R01, t01, xc0, xc1, max_theta, beta_thresh):
    """L1Angular, helpers.cpp:157-215.  Returns (ok, X in the first view's camera frame)."""
    t10, m0, m1 = _bearings(R01, t01, xc0, xc1)
    a0 = _f32(np.linalg.norm(np.cross(m0 / np.linalg.norm(m0), t10)))
    a1 = _f32(np.linalg.norm(np.cross(m1 / np.linalg.norm(m1), t10)))
    if a0 <= a1:
        n1 = np.cross(m1, t10)
        n1 = n1 / np.linalg.norm(n1)
        m0p, m1p = m0 - (m0 @ n1) * n1, m1
    else:
        n0 = np.cross(m0, t10)
        n0 = n0 / np.linalg.norm(n0)
        m0p, m1p = m0, m1 - (m1 @ n0) * n0
    return _angular_finish(R01, t01, t10, m0, m1, m0p, m1p, max_theta, beta_thresh)


def tri_l2_angular(R01, t01, xc0, xc1, max_theta, beta_thresh):
    """L2Angular, helpers.cpp:218-275: n' = right singular vector of the second singular value of
    B = [m0^ m1^]^T (I - t^ t^T)  (JacobiSVD, FullV; the sign of n' cancels in the projections)."""
    t10, m0, m1 = _bearings(R01, t01, xc0, xc1)
    A = np.stack([m0 / np.linalg.norm(m0), m1 / np.linalg.norm(m1)], axis=1)
    th = t10 / np.linalg.norm(t10)
    B = A.T @ (np.eye(3) - np.outer(th, th))
    n = np.linalg.svd(B)[2][1]
    m0p = m0 - (m0 @ n) * n
    m1p = m1 - (m1 @ n) * n
    return _angular_finish(R01, t01, t10, m0, m1, m0p, m1p, max_theta, beta_thresh)


def tri_linf_angular(R01, t01, xc0, xc1, max_theta, beta_thresh):
    """LinfAngular, helpers.cpp:277-327.  n' is NOT normalised there (`n_prime_hat` is n_a or n_b as computed); kept."""
    t10, m0, m1 = _bearings(R01, t01, xc0, xc1)
    m0h, m1h = m0 / np.linalg.norm(m0), m1 / np.linalg.norm(m1)
    na, nb = np.cross(m0h + m1h, t10), np.cross(m0h - m1h, t10)
    n = na if np.linalg.norm(na) >= np.linalg.norm(nb) else nb
    m0p = m0 - (m0 @ n) * n
    m1p = m1 - (m1 @ n) * n
    return _angular_finish(R01, t01, t10, m0, m1, m0p, m1p, max_theta, beta_thresh)


def tri_dlt_svd(R12, t12, xc1, xc2):
    """DirectLinearTransformSVD, helpers.cpp:103-129 (always 'succeeds')."""
    P1 = np.zeros((3, 4))
    P1[:, :3] = np.eye(3)
    P2 = np.zeros((3, 4))
    P2[:, :3] = R12.T
    P2[:, 3] = -R12.T @ t12
    f1 = np.array([xc1[0], xc1[1], 1.0])
    f1 = f1 / np.linalg.norm(f1)
    f2 = np.array([xc2[0], xc2[1], 1.0])
    f2 = f2 / np.linalg.norm(f2)
    A = np.stack([f1[0] * P1[2] - f1[2] * P1[0], f1[1] * P1[2] - f1[2] * P1[1], f2[0] * P2[2] - f2[2] * P2[0], f2[1] * P2[2] - f2[2] * P2[1]])
    v = np.linalg.svd(A)[2][3]
    with np.errstate(all="ignore"):
        return True, v[:3] / v[3]


def tri_dlt_avg(R12, t12, xc1, xc2):
    """DirectLinearTransformAvg, helpers.cpp:131-154 (mid-point of the two rays; always 'succeeds')."""
    f1 = np.array([xc1[0], xc1[1], 1.0])
    f1 = f1 / np.linalg.norm(f1)
    f2 = np.array([xc2[0], xc2[1], 1.0])
    f2 = f2 / np.linalg.norm(f2)
    f2u = R12 @ f2
    b = np.array([t12 @ f1, t12 @ f2u])
    A = np.array([[f1 @ f1, -(f1 @ f2u)], [f1 @ f2u, -(f2u @ f2u)]])
    det = A[0, 0] * A[1, 1] - A[0, 1] * A[1, 0]
    with np.errstate(all="ignore"):
        inv = np.array([[A[1, 1], -A[0, 1]], [-A[1, 0], A[0, 0]]]) * (1.0 / det)  # Eigen's 2x2 inverse: adjugate * (1 / determinant)
        lam = inv @ b
    return True, (lam[0] * f1 + (t12 + lam[1] * f2u)) / 2.0


TRI_METHODS = ("direct_linear_transform_svd", "direct_linear_transform_avg", "l1_angular", "l2_angular", "linf_angular")


def triangulate(method, R12, t12, xc1, xc2, zmin, zmax, max_theta, beta_thresh):
    """Feature::Triangulate after the un-projection (feature.cpp:695-749): g12 = (ref.gsb*gbc)^-1 (gsb*gbc), xc1 / xc2 the
    normalised first / newest observation.  Returns the new feature state [x/z, y/z, log z] or None (bad triangulation)."""
    if method == "direct_linear_transform_svd":
        ok, X = tri_dlt_svd(R12, t12, xc1, xc2)
    elif method == "direct_linear_transform_avg":
        ok, X = tri_dlt_avg(R12, t12, xc1, xc2)
    elif method == "l1_angular":
        ok, X = tri_l1_angular(R12, t12, xc1, xc2, max_theta, beta_thresh)
    elif method == "l2_angular":
        ok, X = tri_l2_angular(R12, t12, xc1, xc2, max_theta, beta_thresh)
    elif method == "linf_angular":
        ok, X = tri_linf_angular(R12, t12, xc1, xc2, max_theta, beta_thresh)
    else:
        raise ValueError("Incorrect Method for Triangulation: " + str(method))
    if not ok:
        return None
    z = X[2]
    if z < zmin or z > zmax:  # NaN passes both comparisons in the reference too (then log(NaN) poisons the feature)
        return None
    return np.array([X[0] / z, X[1] / z, math.log(z) if z > 0 else float("nan")])


def refine_depth(cam: Camera, x, P, gref, gbc, views, ref_index, two_view, use_hessian, max_iters, eps, max_res_norm, Rtri):
    """Feature::RefineDepth (src/feature.cpp:299-420; APPROXIMATE_INIT_COVARIANCE is off in the reference build): Gauss-Newton on the
    feature's local state x = [x/z, y/z, log z] over its stored observations.  views: [(Rsb, Tsb, xp)] of the observing groups in the
    reference's iteration order, ref_index the entry of the reference group (skipped).  Returns (ok, x, P, Xs) — Xs is the cached
    Feature::Xs_ the last evaluation leaves behind.
    Quirks kept: `two_view` picks the first and the last observation (std::minmax_element with a comparator that is always false,
    feature.cpp:305-310); the accept test compares the SUM of the residual norms with `max_res_norm`; after the last iteration the
    state has taken a step whose residual is never evaluated; on a revert the Hessian of the rejected state is the one use_hessian keeps."""
    Rr, Tr = gref
    Rbc, Tbc = gbc
    idx = list(range(len(views)))
    if two_view:
        idx = [idx[0], idx[-1]]
    x = np.asarray(x, dtype=np.float64).copy()
    x0 = x.copy()
    H = np.zeros((3, 3))
    res_norm0 = 0.0
    Xs = None
    Rsc, Tsc = Rr @ Rbc, Rr @ Tbc + Tr
    _err = np.seterr(all="ignore")  # a diverging step yields inf / NaN in the reference too; it is carried, not raised
    for it in range(max_iters):
        Xc, dXc_dx = unproject_logz(x)
        Xs = Rsc @ Xc + Tsc
        dXs_dx = Rsc @ dXc_dx
        H = np.zeros((3, 3))
        b = np.zeros(3)
        res_norm = 0.0
        for k in idx:
            if k == ref_index:
                continue
            Rg, Tg, xp_obs = views[k]
            Rgc, Tgc = Rg @ Rbc, Rg @ Tbc + Tg
            Xcn = Rgc.T @ (Xs - Tgc)
            dXcn_dx = Rgc.T @ dXs_dx
            xcn, dxcn_dXcn = project(Xcn)
            xp, dxp_dxcn = cam.project(xcn)
            J = dxp_dxcn @ dxcn_dXcn @ dXcn_dx
            H += J.T @ J / Rtri
            res = xp - np.asarray(xp_obs, dtype=np.float64)
            b += J.T @ res / Rtri
            res_norm += float(np.linalg.norm(res))
        if it > 0 and res_norm > res_norm0:
            x = x0.copy()  # RestoreState
            break
        # completeOrthogonalDecomposition().solve(b): minimum-norm least squares; a rank-0 decomposition (NaN Hessian, see pinv_sym3)
        # returns the zero vector whatever b holds
        delta = pinv_sym3(H) @ b if np.isfinite(H).all() else np.zeros(3)
        x0 = x.copy()  # BackupState
        x = x - delta
        res_norm0 = res_norm
        if np.abs(delta).max() < eps:
            break
    np.seterr(**_err)
    if res_norm0 > max_res_norm:
        return False, x, P, Xs
    if use_hessian:
        Hp = pinv_sym3(H)
        if np.isnan(Hp).any():
            return False, x, P, Xs
        P = Hp
    return True, x, P, Xs


def pinv_sym3(H):
    """Pseudo-inverse of the (symmetric, PSD) 3x3 Gauss-Newton Hessian with Eigen's rank rule of completeOrthogonalDecomposition: pivots
    below epsilon * 3 * max pivot count as zero (rank 2 when a single view constrains the feature)."""
    if not np.isfinite(H).all():
        # a diverged iterate (log-depth of several hundred: exp overflows, inf - inf = NaN) makes every pivot NaN; Eigen's rank() counts
        # pivots with `abs(pivot) > threshold`, which is false for NaN: rank 0, and solve() / pseudoInverse() of a rank-0 decomposition
        # return ZERO (CompleteOrthogonalDecomposition::_solve_impl).  So the step is 0, the loop ends, and use_hessian stores P = 0.
        return np.zeros((3, 3))
    w, V = np.linalg.eigh(0.5 * (H + H.T))
    wmax = np.abs(w).max()
    out = np.zeros((3, 3))
    for i in range(3):
        if abs(w[i]) > 3 * np.finfo(float).eps * wmax and wmax > 0:
            out += np.outer(V[:, i], V[:, i]) / w[i]
    return out


def predict_pixel(cam: Camera, x, gref, gsb, gbc):
    """Feature::Predict, src/feature.h:175-179 with Feature::Xs (feature.cpp:108-118)."""
    Xc, _ = unproject_logz(x)
    Rr, Tr = gref
    Rbc, Tbc = gbc
    Rsb, Tsb = gsb
    Xs = Rr @ (Rbc @ Xc + Tbc) + Tr
    Rsc, Tsc = Rsb @ Rbc, Rsb @ Tbc + Tsb
    Xcn = Rsc.T @ (Xs - Tsc)
    xcn, _ = project(Xcn)
    return cam.project(xcn)[0]


# ----------------------------------------------------------------------------
# OOS / MSCKF Jacobian + nullspace projection (dead code at runtime in the reference)
# ----------------------------------------------------------------------------
def oos_jacobian(lay: Layout, cam: Camera, Xs, obs, Rbc, Tbc):
    """Feature::ComputeOOSJacobianInternal, src/oos.cpp:39-89, for a list of
    observations obs = [(Rsb_g, Tsb_g, g_sind, xp_meas)].  Returns Hf (2k x 3),
    Hx (2k x N), inn (2k)."""
    k = len(obs)
    Hf = np.zeros((2 * k, 3))
    Hx = np.zeros((2 * k, lay.N))
    inn = np.zeros(2 * k)
    Rbc_t = Rbc.T
    for i, (Rsb, Tsb, gs, xpm) in enumerate(obs):
        Rsb_t = Rsb.T
        Xb = Rsb_t @ (Xs - Tsb)
        Xcn = Rbc_t @ (Xb - Tbc)
        xcn, dxcn_dXcn = project(Xcn)
        xp, dxp_dxcn = cam.project(xcn)
        d = dxp_dxcn @ dxcn_dXcn
        goff = lay.goff(gs)
        inn[2 * i : 2 * i + 2] = np.asarray(xpm) - xp
        Hf[2 * i : 2 * i + 2] = d @ Rbc_t @ Rsb_t
        Hx[2 * i : 2 * i + 2, goff : goff + 3] = d @ Rbc_t @ hat(Xb)
        Hx[2 * i : 2 * i + 2, goff + 3 : goff + 6] = d @ Rbc_t @ (-Rsb_t)
        Hx[2 * i : 2 * i + 2, WBC : WBC + 3] = d @ hat(Xcn)
        Hx[2 * i : 2 * i + 2, TBC : TBC + 3] = d @ (-Rbc_t)
    return Hf, Hx, inn


def givens_pair(a, b, eps=1e-10):
    """static givens(), src/helpers.cpp:27-46 -> (c, s) with G = [[c, s], [-s, c]]."""
    if abs(b) < eps:
        return 1.0, 0.0
    if abs(b) > abs(a):
        t = -a / b
        s = 1 / math.sqrt(1 + t * t)
        return s * t, s
    t = -b / a
    c = 1 / math.sqrt(1 + t * t)
    return c, c * t


def left_nullspace_project(Hf, Hx, inn):
    """What SlowGivens (helpers.cpp:13-23) + oos.cpp:29-30 compute, up to the choice
    of basis: A spans ker(Hf^T); Hx <- A^T Hx, inn <- A^T inn.  The reference's basis
    (FullPivLU::kernel) is neither unique nor orthonormal, so parity is on the
    invariants only: A^T Hf = 0 and rowspace(A^T Hx).  Here A is orthonormal (SVD)."""
    U, s, _ = np.linalg.svd(Hf, full_matrices=True)
    rank = int((s > 1e-12 * max(1.0, s[0])).sum())
    A = U[:, rank:]
    return A, A.T @ Hx, A.T @ inn


# ----------------------------------------------------------------------------
# covariance slot surgery
# ----------------------------------------------------------------------------
def add_group_to_state(lay: Layout, P, err, sind):
    """Estimator::AddGroupToState, src/estimator.cpp:786-823 (sequential row then
    column copies — the order matters for the diagonal block)."""
    off = lay.goff(sind)
    err[off : off + 3] = err[WSB : WSB + 3]
    err[off + 3 : off + 6] = err[TSB : TSB + 3]
    P[off : off + 3, :] = P[WSB : WSB + 3, :]
    P[:, off : off + 3] = P[:, WSB : WSB + 3]
    P[off + 3 : off + 6, :] = P[TSB : TSB + 3, :]
    P[:, off + 3 : off + 6] = P[:, TSB : TSB + 3]


def remove_group_from_state(lay: Layout, P, err, sind):
    """Estimator::RemoveGroupFromState, src/estimator.cpp:739-761."""
    off = lay.goff(sind)
    err[off : off + 6] = 0
    P[off : off + 6, :] = 0
    P[:, off : off + 6] = 0


def add_feature_to_state(lay: Layout, P, sind, Pf):
    """Estimator::AddFeatureToState + Feature::FillCovarianceBlock,
    src/estimator.cpp:825-846, src/feature.cpp:753-776."""
    off = lay.foff(sind)
    P[off : off + 3, :] = 0
    P[:, off : off + 3] = 0
    P[off : off + 3, off : off + 3] = Pf


def remove_feature_from_state(lay: Layout, P, err, sind):
    """Estimator::RemoveFeatureFromState, src/estimator.cpp:763-784."""
    off = lay.foff(sind)
    err[off : off + 3] = 0
    P[off : off + 3, :] = 0
    P[:, off : off + 3] = 0


def fix_feature_xy(lay: Layout, P, sind):
    """Estimator::FixFeatureXY, src/estimator.cpp:1474-1478."""
    off = lay.foff(sind)
    P[off : off + 2, :] = 0
    P[:, off : off + 2] = 0


def switch_ref_group_cov(lay: Layout, P, sind, degrees_fixed):
    """covariance part of Estimator::SwitchRefGroup, src/estimator.cpp:1379-1390."""
    off = lay.goff(sind)
    if degrees_fixed == 4:
        P[off + 2 : off + 6, :] = 0
        P[:, off + 2 : off + 6] = 0
    else:
        P[off : off + 6, :] = 0
        P[:, off : off + 6] = 0


# ----------------------------------------------------------------------------
# IMU propagation ("next" row f1): nominal state + covariance
# ----------------------------------------------------------------------------
@dataclasses.dataclass
class MotionState:
    Rsb: np.ndarray
    Tsb: np.ndarray
    Vsb: np.ndarray
    bg: np.ndarray
    ba: np.ndarray
    Rbc: np.ndarray
    Tbc: np.ndarray
    Rsg: np.ndarray
    counter: int = 0

    def copy(self):
        return MotionState(*(np.array(v, copy=True) if isinstance(v, np.ndarray) else v for v in dataclasses.astuple(self)))


def quat_normalize_rot(R):
    """SO3::normalize() (Sophus normalises the unit quaternion): project to SO(3)."""
    U, _, Vt = np.linalg.svd(R)
    return U @ Vt


def compose_motion(X: MotionState, V, gyro, accel, dt, Cg, Ca, g):
    """Estimator::ComposeMotion, src/estimator.cpp:597-612."""
    gyro_c = Cg @ gyro - X.bg
    accel_c = Ca @ accel - X.ba
    X.Tsb = X.Tsb + V * dt
    X.Vsb = X.Vsb + (X.Rsb @ accel_c + X.Rsg @ g) * dt
    X.Rsb = X.Rsb @ so3_exp(gyro_c * dt)


def motion_jacobian(X: MotionState, gyro, accel, Cg, Ca, g, rec=None, h=0.0):
    """Estimator::ComputeMotionJacobianAt, src/estimator.cpp:614-702 -> F (23x23), G (23x12).
    rec (optional list) receives what F, G depend on: (Rsb, calibrated gyro, calibrated accel, h)."""
    gyro_c = Cg @ gyro - X.bg
    accel_c = Ca @ accel - X.ba
    if rec is not None:
        rec.append(np.concatenate([X.Rsb.ravel(), gyro_c, accel_c, [h]]))
    R = X.Rsb
    F = np.zeros((K_MOTION, K_MOTION))
    G = np.zeros((K_MOTION, 12))
    F[WSB : WSB + 3, WSB : WSB + 3] = -hat(gyro_c)
    F[WSB : WSB + 3, BG : BG + 3] = -np.eye(3)
    F[TSB : TSB + 3, VSB : VSB + 3] = np.eye(3)
    F[VSB : VSB + 3, WSB : WSB + 3] = -R @ hat(accel_c)
    F[VSB : VSB + 3, BA : BA + 3] = -R
    F[VSB : VSB + 3, WSG : WSG + 2] = (-R @ hat(g))[:, :2]
    G[WSB : WSB + 3, 0:3] = -np.eye(3)
    G[BG : BG + 3, 6:9] = np.eye(3)
    G[BA : BA + 3, 9:12] = np.eye(3)
    G[VSB : VSB + 3, 3:6] = -R
    return F, G


_PD_A = [
    (2.0 / 9.0, [2.0 / 9.0]),
    (3.0 / 9.0, [1.0 / 12.0, 3.0 / 12.0]),
    (5.0 / 9.0, [55.0 / 324.0, -75.0 / 324.0, 200.0 / 324.0]),
    (6.0 / 9.0, [83.0 / 330.0, -195.0 / 330.0, 305.0 / 330.0, 27.0 / 330.0]),
    (1.0, [-19.0 / 28.0, 63.0 / 28.0, 4.0 / 28.0, -108.0 / 28.0, 88.0 / 28.0]),
    (1.0, [38.0 / 400.0, 0.0, 240.0 / 400.0, -243.0 / 400.0, 330.0 / 400.0, 35.0 / 400.0]),
]
_PD_B = [0.0862, 0.0, 0.6660, -0.7857, 0.9570, 0.0965, -0.0200]


def prince_dormand_step(X: MotionState, Pmm, gyro0, accel0, slope_gyro, slope_accel, dt, Cg, Ca, g, Qimu, rec=None, h_enc=None):
    """Estimator::PrinceDormandStep, src/princedormand.cpp:85-221.  Returns
    (F_total 23x23, new Pmm); mutates X.  The strip update P[0:23,23:] = F P[0:23,23:]
    (:211-215) is applied by the caller."""
    Ks, FKs, PKs = [], [], []
    F, G = motion_jacobian(X, gyro0, accel0, Cg, Ca, g, rec, dt if h_enc is None else h_enc)
    GQG = lambda G_: G_ @ Qimu @ G_.T
    Ks.append(X.Vsb.copy())
    FKs.append(F.copy())
    PKs.append(F @ Pmm + Pmm @ F.T + GQG(G))
    for c, a in _PD_A:
        X0 = X.copy()
        step = c * dt
        gy = gyro0 + slope_gyro * step
        ac = accel0 + slope_accel * step
        V = sum(ai * Ki for ai, Ki in zip(a, Ks))
        compose_motion(X0, V, gy, ac, step, Cg, Ca, g)
        X0.Rsb = quat_normalize_rot(X0.Rsb)
        F, G = motion_jacobian(X0, gy, ac, Cg, Ca, g, rec)
        Ks.append(X0.Vsb.copy())
        FKs.append(F + F @ sum(ai * FKi for ai, FKi in zip(a, FKs)) * dt)
        P0 = Pmm + sum(ai * PKi for ai, PKi in zip(a, PKs)) * dt
        PKs.append(F @ P0 + P0 @ F.T + GQG(G))
    K = sum(b * Ki for b, Ki in zip(_PD_B, Ks))
    FK = sum(b * FKi for b, FKi in zip(_PD_B, FKs))
    PK = sum(b * PKi for b, PKi in zip(_PD_B, PKs))
    compose_motion(X, K, gyro0 + slope_gyro * dt, accel0 + slope_accel * dt, dt, Cg, Ca, g)
    X.Rsb = quat_normalize_rot(X.Rsb)
    Ftot = np.eye(K_MOTION) + FK * dt
    return Ftot, Pmm + PK * dt


def rk4_step(X: MotionState, Pmm, gyro0, accel0, slope_gyro, slope_accel, dt, Cg, Ca, g, Qimu, rec=None, h_enc=None):
    """Estimator::RK4Step, src/rk4.cpp:35-103 (including its use of the half-step
    input for the 4th stage, :79)."""
    half = 0.5 * dt
    GQG = lambda G_: G_ @ Qimu @ G_.T
    F, G = motion_jacobian(X, gyro0, accel0, Cg, Ca, g, rec, dt if h_enc is None else h_enc)
    K1, FK1 = X.Vsb.copy(), F.copy()
    PK1 = F @ Pmm + Pmm @ F.T + GQG(G)
    gy, ac = gyro0 + half * slope_gyro, accel0 + half * slope_accel
    X0 = X.copy()
    compose_motion(X0, 0.5 * K1, gy, ac, half, Cg, Ca, g)
    X0.Rsb = quat_normalize_rot(X0.Rsb)
    K2 = X0.Vsb.copy()
    F, G = motion_jacobian(X0, gy, ac, Cg, Ca, g, rec)
    FK2 = F + F @ FK1 * half
    P0 = Pmm + half * PK1
    PK2 = F @ P0 + P0 @ F.T + GQG(G)
    X0 = X.copy()
    compose_motion(X0, 0.5 * K2, gy, ac, half, Cg, Ca, g)
    X0.Rsb = quat_normalize_rot(X0.Rsb)
    K3 = X0.Vsb.copy()
    F, G = motion_jacobian(X0, gy, ac, Cg, Ca, g, rec)
    FK3 = F + F @ FK2 * half
    P0 = Pmm + half * PK2
    PK3 = F @ P0 + P0 @ F.T + GQG(G)
    X0 = X.copy()
    compose_motion(X0, K3, gy, ac, dt, Cg, Ca, g)
    X0.Rsb = quat_normalize_rot(X0.Rsb)
    K4 = X0.Vsb.copy()
    F, G = motion_jacobian(X0, gy, ac, Cg, Ca, g, rec)
    FK4 = F + F @ FK3 * dt
    P0 = Pmm + dt * PK3
    PK4 = F @ P0 + P0 @ F.T + GQG(G)
    Kt = (K1 + 2.0 * (K2 + K3) + K4) / 6.0
    FK = (FK1 + 2.0 * (FK2 + FK3) + FK4) / 6.0
    PK = (PK1 + 2.0 * (PK2 + PK3) + PK4) / 6.0
    compose_motion(X, Kt, gyro0 + dt * slope_gyro, accel0 + dt * slope_accel, dt, Cg, Ca, g)
    X.Rsb = quat_normalize_rot(X.Rsb)
    return np.eye(K_MOTION) + FK * dt, Pmm + PK * dt


def integrate(method, X, Pmm, gyro0, accel0, slope_gyro, slope_accel, dt, Cg, Ca, g, Qimu, h0=0.002, rec=None):
    """Fixed-step driver with the half-step trick: Estimator::PrinceDormand
    (princedormand.cpp:60-81, control_stepsize=false) / Estimator::RK4 (rk4.cpp:14-31).
    Returns (Phi = product of per-substep F, Pmm)."""
    step_fn = prince_dormand_step if method == "PrinceDormand" else rk4_step
    Phi = np.eye(K_MOTION)
    if h0 < 0:
        F, Pmm = step_fn(X, Pmm, gyro0, accel0, slope_gyro, slope_accel, dt, Cg, Ca, g, Qimu, rec, -dt)
        return F @ Phi, Pmm
    total = 0.0
    gyro, accel = gyro0.copy(), accel0.copy()
    while total < dt:
        h = h0
        if total + h > dt:
            h = dt - total
        elif total + h + 0.5 * h > dt:
            h = 0.5 * h
        last = not (total + h < dt)
        F, Pmm = step_fn(X, Pmm, gyro, accel, slope_gyro, slope_accel, h, Cg, Ca, g, Qimu, rec, -h if last else h)
        Phi = F @ Phi
        gyro = gyro + slope_gyro * h
        accel = accel + slope_accel * h
        total += h
    return Phi, Pmm


def apply_propagation(P, Phi, Pmm_new, Qmodel):
    """Covariance side of Propagate: motion block replaced by the integrated block
    (+Qmodel, src/estimator.cpp:590) and cross strips P[0:23,23:] <- Phi P[0:23,23:]
    (princedormand.cpp:211-215 composed over the substeps)."""
    m = K_MOTION
    P[:m, :m] = Pmm_new + Qmodel
    P[:m, m:] = Phi @ P[:m, m:]
    P[m:, :m] = P[m:, :m] @ Phi.T
