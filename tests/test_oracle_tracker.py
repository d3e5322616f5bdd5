"""Pins oracle/tracker_oracle.c against (a) the committed cv2 golden vectors and (b) cv2 itself
when it is importable (it is in this image).  CPU only."""
import os

import numpy as np
import pytest

from oracle import tracker_oracle as T
from xivo_b200 import synth

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "tracker_cv2.npz"))


def test_pyrdown_matches_golden():
    lv = T.pyramid(G["a"], 15, 5)
    assert len(lv) == 3
    for i, l in enumerate(lv):
        assert np.array_equal(l, G[f"pyr{i}"])
    lv3 = T.pyramid(G["pyr3_0"], 15, 5)
    for i, l in enumerate(lv3):
        assert np.array_equal(l, G[f"pyr3_{i}"])


def test_gray_matches_golden():
    assert np.array_equal(T.bgr2gray(G["pyr3_0"]), G["gray_from_bgr"])


@pytest.mark.parametrize("thr", [10, 20])
def test_fast_matches_golden(thr):
    xy, sc, n = T.fast_detect(G["a"], thr)
    g = G[f"fast{thr}"]
    assert n == len(g)
    assert np.array_equal(xy, g[:, :2]) and np.array_equal(sc, g[:, 2])


def test_fast_bgr_matches_golden():
    xy, sc, n = T.fast_detect(G["pyr3_0"], 20)
    g = G["fast20_bgr"]
    assert n == len(g) and np.array_equal(xy, g[:, :2]) and np.array_equal(sc, g[:, 2])


@pytest.mark.parametrize("name,img", [("gray", ("a", "b")), ("bgr", None)])
def test_lk_matches_golden(name, img):
    if name == "gray":
        a, b = G["a"], G["b"]
    else:
        a, b = synth.to_bgr(G["a"], True), synth.to_bgr(G["b"], True)
    p1, st, er = T.lk_track(a, b, G["lk_p0"], G["lk_init"])
    assert np.array_equal(st, G[f"lk_{name}_st"])
    ok = st == 1
    # OpenCV sums float SIMD lanes; the oracle sums exactly -> sub-milli-pixel differences
    assert np.abs(p1[ok] - G[f"lk_{name}_p1"][ok]).max() < 5e-3
    assert np.abs(er[ok] - G[f"lk_{name}_err"][ok]).max() < 5e-3


def test_against_live_cv2_full_size():
    cv2 = pytest.importorskip("cv2")
    a, b = synth.frame_pair(480, 640, seed=0)
    assert np.array_equal(T.pyrdown(a), cv2.pyrDown(a))
    n, pyr = cv2.buildOpticalFlowPyramid(a, (15, 15), 5, withDerivatives=True)
    assert n + 1 == len(T.pyramid(a, 15, 5))
    assert np.array_equal(T.scharr(a), pyr[1])
    kps = cv2.FastFeatureDetector_create(20, True).detect(a, None)
    xy, sc, n = T.fast_detect(a, 20)
    ref = np.array([[k.pt[0], k.pt[1], k.response] for k in kps], np.int32)
    assert n == len(ref) and np.array_equal(xy, ref[:, :2]) and np.array_equal(sc, ref[:, 2])
    order = np.lexsort((xy[:, 0], xy[:, 1], -sc))[:200]
    p0 = xy[order].astype(np.float32)
    crit = (cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, 30, 0.01)
    p1c, stc, _ = cv2.calcOpticalFlowPyrLK(a, b, p0, p0.copy(), winSize=(15, 15), maxLevel=5, criteria=crit,
                                           flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    p1, st, _ = T.lk_track(a, b, p0, p0)
    assert (st == stc.ravel()).mean() > 0.99
    ok = (st == 1) & (stc.ravel() == 1)
    assert np.abs(p1[ok] - p1c[ok]).max() < 5e-3
    assert np.abs((p1[ok] - p0[ok]).mean(0) - np.array([-3.0, -2.0])).max() < 0.1


def test_selection_logic_spacing():
    a, _ = synth.frame_pair(240, 320, seed=5)
    xy, sc, _ = T.fast_detect(a, 20)
    m = T.Mask(240, 320, margin=8, mask_size=15)
    m.reset()
    picked = T.select_keypoints(m, xy, sc, 60)
    assert 0 < len(picked) <= 60
    pts = xy[picked]
    assert pts[:, 0].min() >= 8 and pts[:, 0].max() < 320 - 8 and pts[:, 1].min() >= 8 and pts[:, 1].max() < 240 - 8
    d = np.abs(pts[:, None, :] - pts[None, :, :]).max(-1) + np.eye(len(pts)) * 100
    assert d.min() > 7  # no two picks inside each other's 15x15 box
    assert all(sc[picked[i]] >= sc[picked[i + 1]] for i in range(len(picked) - 1))


def test_homography_sampling_kernel_and_ransac_loop_match_cv2():
    """oracle/homography_oracle.py (work in progress for SURVEY §8f row 3): the parts that are pinned on cv2 4.13.  On unrelated point
    sets every 4-point hypothesis has its own inlier set (the sample itself plus accidents), so identical masks mean the same cv::RNG
    draws, the same checkSubset decisions, the same 4-point kernel and the same adaptive stopping rule."""
    cv2 = pytest.importorskip("cv2")
    from oracle import homography_oracle as HO

    rng = np.random.default_rng(1)
    for _ in range(3):
        p0 = rng.uniform(0, 500, (4, 2)).astype(np.float32)
        p1 = (p0 + rng.normal(0, 5, (4, 2))).astype(np.float32)
        Hc, _m = cv2.findHomography(p0, p1, 0)
        assert np.abs(Hc - HO.run_kernel(p0, p1)).max() <= 1e-10 * np.abs(Hc).max()
    for _ in range(40):
        n = int(rng.integers(5, 40))
        p0 = rng.uniform(0, 500, (n, 2)).astype(np.float32)
        p1 = rng.uniform(0, 500, (n, 2)).astype(np.float32)
        _H, mc = cv2.findHomography(p0, p1, cv2.RANSAC, 3.0, maxIters=200, confidence=0.995)
        ok, mo = HO.find_homography_mask(p0, p1, HO.RANSAC, 3.0, 200, 0.995)
        assert np.array_equal(np.zeros(n, np.uint8) if mc is None else mc.ravel(), mo)


def test_cross_checked_hamming_matcher_matches_cv2_bfmatcher():
    """orc_bf_match_crosscheck against cv2.BFMatcher(NORM_HAMMING, crossCheck=True).knnMatch(q, t, 1): same (query, train, distance)
    triples, also with few distinct descriptor values (ties: the first index wins on both sides of the cross-check) -- the matcher the
    reference calls at src/tracker.cpp:261-262 and :378-379.  (The descriptor VALUES of the BRIEF restatement are parity-unpinned:
    opencv_contrib's test table is not available; scripts/make_brief_pattern.py.)"""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for trial in range(120):
        nq, nt, nb = int(rng.integers(1, 40)), int(rng.integers(1, 60)), int(rng.integers(1, 5))
        q = rng.integers(0, 256, (nq, nb), dtype=np.uint8)
        t = rng.integers(0, 256, (nt, nb), dtype=np.uint8)
        if trial % 3 == 0:
            q &= 0x03
            t &= 0x03
        m = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).knnMatch(q, t, k=1)
        ref = [(x[0].queryIdx, x[0].trainIdx, int(x[0].distance)) for x in m if len(x)]
        assert T.bf_match_crosscheck(q, t) == ref, f"trial {trial}"


def test_brief_restatement_properties():
    """Own-table BRIEF-32: border rule of KeyPointsFilter::runByImageBorder (28 px), (int)(pt + 0.5) rounding, bit order, and the box sums
    against a numpy integral image (what opencv_contrib reads its smoothed values from)."""
    from xivo_b200 import synth

    a, _ = synth.frame_pair(200, 260, seed=3)
    kp = np.array([[100.2, 120.7], [10, 10], [231.9, 50], [232.0, 50], [28, 28], [27.99, 100], [100.49, 120.5], [100.5, 120.49]], np.float32)
    d, v = T.brief(a, kp)
    assert v.tolist() == [True, False, True, False, True, False, True, True]
    assert not d[1].any() and d[0].any()
    assert np.array_equal(d[6], T.brief(a, np.array([[100, 121]], np.float32))[0][0])  # (int)(x + .5): 100.49 -> 100, 120.5 -> 121
    assert np.array_equal(d[7], T.brief(a, np.array([[101, 120]], np.float32))[0][0])
    # independent evaluation through an integral image
    import re
    pat = np.array([[int(x) for x in m] for m in re.findall(r"\{(-?\d+), (-?\d+), (-?\d+), (-?\d+)\}", open(os.path.join(os.path.dirname(__file__), "..", "xivo_b200", "csrc", "brief_pattern.h")).read())])
    assert pat.shape == (256, 4) and np.abs(pat).max() <= 19
    S = np.zeros((201, 261), np.int64)
    S[1:, 1:] = a.astype(np.int64).cumsum(0).cumsum(1)
    box = lambda x, y: S[y + 5, x + 5] - S[y - 4, x + 5] - S[y + 5, x - 4] + S[y - 4, x - 4]
    cx, cy = 100, 121
    bits = [int(box(cx + p[0], cy + p[1]) < box(cx + p[2], cy + p[3])) for p in pat]
    want = np.packbits(np.array(bits, np.uint8))  # MSB first = test 8 i + k -> bit 7 - k
    assert np.array_equal(d[0], want)
