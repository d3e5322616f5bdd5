"""Restatement of cv::findHomography's inlier mask (LMEDS and RANSAC) as called by the reference's tracker-level outlier rejection,
Tracker::OutlierRejection (/root/reference/src/tracker.cpp:705-753, options :131-150).  TEST INFRASTRUCTURE ONLY; the product's version
is xivo_b200/csrc/homography.h.

The arithmetic lives in OpenCV (un-vendored, version unpinned in the reference): modules/calib3d/src/fundam.cpp
(HomographyEstimatorCallback::{checkSubset, runKernel, computeError}, HomographyRefineCallback, findHomography), ptsetreg.cpp
({RANSAC,LMeDS}PointSetRegistrator, RANSACUpdateNumIters), levmarq.cpp (LMSolver) and core's cv::RNG.  Pinned on cv2 4.13 in this
container (tests/test_oracle_tracker.py, tests/test_host_logic.py): the cv::RNG sequence, getSubset / checkSubset, the 4-point normalised
DLT kernel (1e-13 of cv2 with method 0), the RANSAC / LMedS loops and `find_homography_mask_413` = the mask cv2 returns (identical on
300 structured scenes; H within 1e-6).

A version finding: cv2 4.13 does NOT return the estimator's mask.  After re-estimating H on the estimator's inliers and refining it with
Levenberg-Marquardt (10 iterations) it returns `computeError(H_refined) <= reprojThreshold^2` for RANSAC *and* LMEDS, i.e. LMEDS's own
sigma = 2.5*1.4826*(1+5/(n-4))*sqrt(median) threshold no longer decides the mask.  OpenCV 3.4 — the version the reference's build evidence
points to — returns the estimator's mask (`find_homography_mask` below).  The reference does not pin a version
(`find_package(OpenCV REQUIRED)`); oracle and product follow the version that can be run and checked here, like the rest of the tracker.

Points are float32 pairs (cv::Point2f); the model is estimated in double and applied in float exactly as computeError does."""
from __future__ import annotations

import math

import numpy as np

LMEDS, RANSAC = 4, 8  # cv::LMEDS, cv::RANSAC
_f32 = np.float32
FLT_EPSILON, DBL_EPSILON, DBL_MIN = 1.1920929e-07, 2.220446049250313e-16, 2.2250738585072014e-308


class CvRNG:
    """cv::RNG: multiply-with-carry, state = (uint32)state * 4164903690 + (state >> 32)."""

    def __init__(self, state=0xFFFFFFFFFFFFFFFF):
        self.state = state if state else 0xFFFFFFFF

    def next(self):
        self.state = ((self.state & 0xFFFFFFFF) * 4164903690 + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a, b):
        return a if a == b else self.next() % (b - a) + a


def update_num_iters(p, ep, model_points, max_iters):
    """RANSACUpdateNumIters (ptsetreg.cpp)."""
    p, ep = min(max(p, 0.0), 1.0), min(max(ep, 0.0), 1.0)
    num = max(1.0 - p, DBL_MIN)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < DBL_MIN:
        return 0
    num, denom = math.log(num), math.log(denom)
    if denom >= 0 or -num >= max_iters * (-denom):
        return max_iters
    return int(np.rint(num / denom))  # cvRound: round half to even


def have_collinear_points(m, count):
    """haveCollinearPoints (fundam.cpp): only the LAST selected point is tested against the earlier pairs."""
    i = count - 1
    for j in range(i):
        dx1, dy1 = float(m[j][0]) - float(m[i][0]), float(m[j][1]) - float(m[i][1])
        for k in range(j):
            dx2, dy2 = float(m[k][0]) - float(m[i][0]), float(m[k][1]) - float(m[i][1])
            if abs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (abs(dx1) + abs(dy1) + abs(dx2) + abs(dy2)):
                return True
    return False


def _det3(a):
    return (a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0])
            + a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]))


def check_subset(ms1, ms2, count):
    """HomographyEstimatorCallback::checkSubset: collinearity, then the orientation of every triple must be preserved (or reversed) by all four."""
    if have_collinear_points(ms1, count) or have_collinear_points(ms2, count):
        return False
    if count == 4:
        negative = 0
        for t in ((0, 1, 2), (1, 2, 3), (0, 2, 3), (0, 1, 3)):
            A = [[float(ms1[i][0]), float(ms1[i][1]), 1.0] for i in t]
            B = [[float(ms2[i][0]), float(ms2[i][1]), 1.0] for i in t]
            negative += _det3(A) * _det3(B) < 0
        if negative != 0 and negative != 4:
            return False
    return True


def run_kernel(M, m):
    """HomographyEstimatorCallback::runKernel: normalised DLT, H = eigenvector of the smallest eigenvalue of L^T L.  Returns H (3x3) or None."""
    M, m = np.asarray(M, dtype=np.float64), np.asarray(m, dtype=np.float64)
    count = len(M)
    cM, cm = M.sum(0) / count, m.sum(0) / count
    sM, sm = np.abs(M - cM).sum(0), np.abs(m - cm).sum(0)
    if min(abs(sm[0]), abs(sm[1]), abs(sM[0]), abs(sM[1])) < DBL_EPSILON:
        return None
    sm, sM = count / sm, count / sM
    inv_hnorm = np.array([[1.0 / sm[0], 0, cm[0]], [0, 1.0 / sm[1], cm[1]], [0, 0, 1]])
    hnorm2 = np.array([[sM[0], 0, -cM[0] * sM[0]], [0, sM[1], -cM[1] * sM[1]], [0, 0, 1]])
    LtL = np.zeros((9, 9))
    for i in range(count):
        x, y = (m[i][0] - cm[0]) * sm[0], (m[i][1] - cm[1]) * sm[1]
        X, Y = (M[i][0] - cM[0]) * sM[0], (M[i][1] - cM[1]) * sM[1]
        Lx = np.array([X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x])
        Ly = np.array([0, 0, 0, X, Y, 1, -y * X, -y * Y, -y])
        LtL += np.outer(Lx, Lx) + np.outer(Ly, Ly)
    w, V = np.linalg.eigh(LtL)
    H0 = V[:, 0].reshape(3, 3)  # smallest eigenvalue (cv::eigen orders descending and takes row 8)
    H = inv_hnorm @ H0 @ hnorm2
    with np.errstate(all="ignore"):
        return H * (1.0 / H[2, 2])


def compute_error(M, m, H):
    """computeError: the model converted to float, everything in float32."""
    Hf = np.asarray(H, dtype=np.float64).ravel().astype(_f32)
    M, m = np.asarray(M, dtype=_f32), np.asarray(m, dtype=_f32)
    with np.errstate(all="ignore"):
        ww = _f32(1.0) / (Hf[6] * M[:, 0] + Hf[7] * M[:, 1] + _f32(1.0))
        dx = (Hf[0] * M[:, 0] + Hf[1] * M[:, 1] + Hf[2]) * ww - m[:, 0]
        dy = (Hf[3] * M[:, 0] + Hf[4] * M[:, 1] + Hf[5]) * ww - m[:, 1]
        return (dx * dx + dy * dy).astype(_f32)


def get_subset(m1, m2, rng, model_points=4, max_attempts=1000):
    """RANSACPointSetRegistrator::getSubset (OpenCV 4.x): draw distinct indices, accept the first subset that passes checkSubset."""
    count = len(m1)
    for _ in range(max_attempts):
        idx = []
        for i in range(model_points):
            k = rng.uniform(0, count)
            while k in idx:
                k = rng.uniform(0, count)
            idx.append(k)
        ms1, ms2 = m1[idx], m2[idx]
        if check_subset(ms1, ms2, model_points):
            return ms1, ms2
    return None


def find_homography_mask(pts0, pts1, method=LMEDS, reproj_thresh=3.0, max_iters=2000, confidence=0.995):
    """Mask of cv::findHomography(pts0, pts1, method, reproj_thresh, mask, max_iters, confidence).
    Returns (ok, mask uint8[n]); on failure the mask is all zero (fundam.cpp: `tempMask = Mat::zeros`)."""
    m1, m2 = np.asarray(pts0, dtype=_f32).reshape(-1, 2), np.asarray(pts1, dtype=_f32).reshape(-1, 2)
    n = len(m1)
    if reproj_thresh <= 0:
        reproj_thresh = 3.0
    if n < 4:
        return False, np.zeros(n, np.uint8)  # (cv2 raises here; Tracker::OutlierRejection never calls it with < 4 points)
    if n == 4 or method == 0:
        ok = run_kernel(m1, m2) is not None
        return ok, (np.ones(n, np.uint8) if ok else np.zeros(n, np.uint8))
    rng = CvRNG()
    best, mask = None, np.zeros(n, np.uint8)
    if method == LMEDS:
        niters = max(update_num_iters(confidence, 0.45, 4, max_iters), 3)
        min_median = float("inf")
        for it in range(niters):
            sub = get_subset(m1, m2, rng)
            if sub is None:
                if it == 0:
                    return False, np.zeros(n, np.uint8)
                break
            H = run_kernel(*sub)
            if H is None:
                continue
            err = compute_error(m1, m2, H)
            median = float(np.sort(err.view(np.int32))[n // 2].view(_f32))  # std::nth_element on the float bits (errors are >= 0)
            if median < min_median:
                min_median, best = median, H
        if best is None:
            return False, np.zeros(n, np.uint8)
        sigma = max(2.5 * 1.4826 * (1 + 5.0 / (n - 4)) * math.sqrt(min_median), 0.001)
        t = _f32(sigma * sigma)
        mask = (compute_error(m1, m2, best) <= t).astype(np.uint8)
        return int(mask.sum()) >= 4, (mask if int(mask.sum()) >= 4 else np.zeros(n, np.uint8))
    if method == RANSAC:
        niters, max_good, best_mask = max_iters, 0, None
        t = _f32(reproj_thresh * reproj_thresh)
        it = 0
        while it < niters:
            sub = get_subset(m1, m2, rng, max_attempts=10000)
            if sub is None:
                if it == 0:
                    return False, np.zeros(n, np.uint8)
                break
            H = run_kernel(*sub)
            if H is not None:
                cur = (compute_error(m1, m2, H) <= t).astype(np.uint8)
                good = int(cur.sum())
                if good > max(max_good, 3):
                    best_mask, max_good = cur, good
                    niters = update_num_iters(confidence, (n - good) / n, 4, niters)
            it += 1
        if max_good > 0:
            return True, best_mask
        return False, np.zeros(n, np.uint8)
    raise ValueError("method must be 0, LMEDS or RANSAC")


def _refine_residual(h8, M, m, want_J):
    """HomographyRefineCallback::compute (fundam.cpp): residuals (2n) and Jacobian (2n x 8) of the 8-parameter homography, in double."""
    Mx, My = M[:, 0].astype(np.float64), M[:, 1].astype(np.float64)
    ww = h8[6] * Mx + h8[7] * My + 1.0
    ww = np.where(np.abs(ww) > DBL_EPSILON, 1.0 / np.where(ww == 0, 1.0, ww), 0.0)
    xi = (h8[0] * Mx + h8[1] * My + h8[2]) * ww
    yi = (h8[3] * Mx + h8[4] * My + h8[5]) * ww
    r = np.empty(2 * len(M))
    r[0::2], r[1::2] = xi - m[:, 0], yi - m[:, 1]
    if not want_J:
        return r, None
    J = np.zeros((2 * len(M), 8))
    J[0::2, 0], J[0::2, 1], J[0::2, 2] = Mx * ww, My * ww, ww
    J[0::2, 6], J[0::2, 7] = -Mx * ww * xi, -My * ww * xi
    J[1::2, 3], J[1::2, 4], J[1::2, 5] = Mx * ww, My * ww, ww
    J[1::2, 6], J[1::2, 7] = -Mx * ww * yi, -My * ww * yi
    return r, J


def lm_refine(H, M, m, max_iters=10):
    """cv::LMSolver (calib3d/levmarq.cpp, LMSolverImpl::run) on HomographyRefineCallback, as findHomography calls it (10 iterations,
    epsx = epsf = FLT_EPSILON): Levenberg-Marquardt with Nielsen-style damping control (Rlo = 0.25, Rhi = 0.75)."""
    x = np.asarray(H, dtype=np.float64).ravel()[:8].copy()
    r, J = _refine_residual(x, M, m, True)
    S = float(r @ r)
    A, v = J.T @ J, J.T @ r
    D = np.diag(A).copy()
    lam, lc = 1.0, 0.75
    it = 0
    while True:
        Ap = A + np.diag(lam * D)
        d = np.linalg.lstsq(Ap, v, rcond=None)[0]  # solve(Ap, v, d, DECOMP_EIG)
        xd = x - d
        rd, _ = _refine_residual(xd, M, m, False)
        Sd = float(rd @ rd)
        dS = float(d @ (2 * v - A @ d))
        R = (S - Sd) / (dS if abs(dS) > DBL_EPSILON else 1.0)
        if R > 0.75:
            lam *= 0.5
            if lam < lc:
                lam = 0.0
        elif R < 0.25:
            t = float(d @ v)
            nu = (Sd - S) / (t if abs(t) > DBL_EPSILON else 1.0) + 2
            nu = min(max(nu, 2.0), 10.0)
            if lam == 0:
                Ai = np.linalg.pinv(A)  # invert(A, Ap, DECOMP_EIG)
                lam = lc = 1.0 / max(DBL_EPSILON, float(np.abs(np.diag(Ai)).max()))
                nu *= 0.5
            lam *= nu
        if Sd < S:
            S, x = Sd, xd
            r, J = _refine_residual(x, M, m, True)
            A, v = J.T @ J, J.T @ r
        it += 1
        if not (it < max_iters and np.abs(d).max() >= FLT_EPSILON and np.abs(r).max() >= FLT_EPSILON):
            break
    return np.append(x, 1.0).reshape(3, 3)


def find_homography_mask_413(pts0, pts1, method=LMEDS, reproj_thresh=3.0, max_iters=2000, confidence=0.995):
    """The mask cv2 4.13 returns: the estimator's inliers (find_homography_mask) are compressed, the model is re-estimated on them
    (runKernel) and refined (lm_refine), and the FINAL mask is computeError(H_refined) <= reproj_thresh^2 over all points.
    Returns (ok, mask, H)."""
    m1, m2 = np.asarray(pts0, dtype=_f32).reshape(-1, 2), np.asarray(pts1, dtype=_f32).reshape(-1, 2)
    n = len(m1)
    if reproj_thresh <= 0:
        reproj_thresh = 3.0
    ok, mask = find_homography_mask(m1, m2, method, reproj_thresh, max_iters, confidence)
    if not ok:
        return False, np.zeros(n, np.uint8), None
    if n == 4 or method == 0:
        return True, mask, run_kernel(m1, m2)
    inl = mask.astype(bool)
    H = run_kernel(m1[inl], m2[inl])
    if H is None:
        return False, np.zeros(n, np.uint8), None
    H = lm_refine(H, m1[inl], m2[inl])
    final = (compute_error(m1, m2, H) <= _f32(reproj_thresh * reproj_thresh)).astype(np.uint8)
    return True, final, H


def tracker_outlier_rejection(pts0, pts1, status, method, reproj_thresh, max_iters, confidence):
    """Tracker::OutlierRejection (tracker.cpp:705-753): runs findHomography on the points whose status is non-zero and clears the status of
    the outliers.  Returns (success, number of rejected outliers or None when it returned early, new status)."""
    status = np.asarray(status, dtype=np.uint8).copy()
    valid = np.nonzero(status)[0]
    if len(valid) < 4:
        return False, None, status
    ok, mask = find_homography_mask(np.asarray(pts0, _f32)[valid], np.asarray(pts1, _f32)[valid], method, reproj_thresh, max_iters, confidence)
    status[valid[mask == 0]] = 0
    return True, int((mask == 0).sum()), status


def tracker_outlier_rejection_413(pts0, pts1, status, method, reproj_thresh, max_iters, confidence):
    """Tracker::OutlierRejection with the mask cv2 4.13 returns (find_homography_mask_413)."""
    status = np.asarray(status, dtype=np.uint8).copy()
    valid = np.nonzero(status)[0]
    if len(valid) < 4:
        return False, None, status
    ok, mask, _H = find_homography_mask_413(np.asarray(pts0, _f32)[valid], np.asarray(pts1, _f32)[valid], method, reproj_thresh, max_iters, confidence)
    status[valid[mask == 0]] = 0
    return True, int((mask == 0).sum()), status
