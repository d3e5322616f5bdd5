"""ctypes wrapper of oracle/tracker_oracle.c (TEST INFRASTRUCTURE ONLY) + the tracker's
selection logic restated from /root/reference/src/tracker.cpp:219-329, :463-629, :760-815."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libtracker_oracle.so")
_lib = None


def _vp(a):
    return C.c_void_p(a.ctypes.data)


def build(force=False):
    src = os.path.join(_HERE, "tracker_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, src, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_lk_track.argtypes = [C.c_void_p] * 2 + [C.c_int] * 3 + [C.c_void_p] * 4 + [C.c_int] * 4 + [C.c_double, C.c_int, C.c_double]
    return _lib


def _shape(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    return img, img.shape[0], img.shape[1], (1 if img.ndim == 2 else img.shape[2])


def bgr2gray(img):
    img, r, c, cn = _shape(img)
    out = np.zeros((r, c), np.uint8)
    lib().orc_bgr2gray(_vp(img), r, c, _vp(out))
    return out


def pyrdown(img):
    img, r, c, cn = _shape(img)
    out = np.zeros(((r + 1) // 2, (c + 1) // 2) + ((cn,) if img.ndim == 3 else ()), np.uint8)
    lib().orc_pyrdown(_vp(img), r, c, cn, _vp(out))
    return out


def pyramid(img, win, max_level):
    img, r, c, cn = _shape(img)
    L = lib().orc_pyramid_levels(r, c, win, max_level)
    out = [img]
    for _ in range(L):
        out.append(pyrdown(out[-1]))
    return out


def brief(img, kp_xy):
    """BRIEF-32 at the given keypoints (n x 2 float32) -> (descriptors n x 32 uint8, valid n bool); BGR input is converted to grey first."""
    img, r, c, cn = _shape(img)
    if cn == 3:
        img = bgr2gray(img)
    kp = np.ascontiguousarray(kp_xy, dtype=np.float32).reshape(-1, 2)
    n = len(kp)
    desc, valid = np.zeros((max(n, 1), 32), np.uint8), np.zeros(max(n, 1), np.uint8)
    if n:
        lib().orc_brief(_vp(img), r, c, _vp(kp), n, _vp(desc), _vp(valid))
    return desc[:n], valid[:n].astype(bool)


def hamming(a, b):
    a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
    return int(lib().orc_hamming(_vp(a), _vp(b), a.size))


def bf_match_crosscheck(query, train):
    """cv::BFMatcher(NORM_HAMMING, crossCheck=True).knnMatch(query, train, 1, compactResult=True) -> [(queryIdx, trainIdx, distance)]."""
    q, t = np.ascontiguousarray(query, np.uint8), np.ascontiguousarray(train, np.uint8)
    if len(q) == 0 or len(t) == 0:
        return []
    out = np.zeros((len(q), 3), np.int32)
    m = lib().orc_bf_match_crosscheck(_vp(q), len(q), _vp(t), len(t), q.shape[1], _vp(out))
    return [tuple(int(v) for v in row) for row in out[:m]]


def scharr(img):
    img, r, c, cn = _shape(img)
    out = np.zeros((r, c, cn * 2), np.int16)
    lib().orc_scharr(_vp(img), r, c, cn, _vp(out))
    return out


def fast_detect(img, threshold, nonmax=True, max_kp=1 << 17):
    img, r, c, cn = _shape(img)
    if cn == 3:
        img = bgr2gray(img)
    xy = np.zeros((max_kp, 2), np.int32)
    sc = np.zeros(max_kp, np.int32)
    n = lib().orc_fast_detect(_vp(img), r, c, int(threshold), int(bool(nonmax)), _vp(xy), _vp(sc), max_kp, None)
    k = min(n, max_kp)
    return xy[:k].copy(), sc[:k].copy(), n


def lk_track(prev, nxt, prev_pts, init_pts, win=15, max_level=5, max_iter=30, eps=0.01, use_initial_flow=True, min_eig=1e-4):
    prev, r, c, cn = _shape(prev)
    nxt = np.ascontiguousarray(nxt, dtype=np.uint8)
    p0 = np.ascontiguousarray(prev_pts, dtype=np.float32).reshape(-1, 2)
    p1 = np.array(init_pts, dtype=np.float32).reshape(-1, 2).copy()
    n = len(p0)
    st = np.zeros(n, np.uint8)
    er = np.zeros(n, np.float32)
    lib().orc_lk_track(_vp(prev), _vp(nxt), r, c, cn, _vp(p0), _vp(p1), _vp(st), _vp(er), n,
                       win, max_level, max_iter, float(eps), int(bool(use_initial_flow)), float(min_eig))
    return p1, st, er


# ---------------------------------------------------------------- selection logic (host side)
class Mask:
    """mask_ of the tracker (tracker.cpp:471-488, :760-774).  NB MaskOut's half size is a
    function-local static frozen at the first call (tracker.cpp:763)."""

    def __init__(self, rows, cols, margin, mask_size):
        self.rows, self.cols, self.margin = rows, cols, margin
        self.half = mask_size >> 1
        self.m = np.zeros((rows, cols), np.uint8)

    def reset(self):
        self.m[:] = 0
        mg = self.margin
        self.m[mg : self.rows - mg, mg : self.cols - mg] = 255

    def valid(self, x, y):
        col, row = int(x), int(y)  # static_cast<int> truncation
        if col < 0 or col >= self.cols or row < 0 or row >= self.rows:
            return False
        return bool(self.m[row, col])

    def mask_out(self, x, y):
        # cv::rectangle(mask, Point2d(x-h, y-h), Point2d(x+h, y+h), 0, FILLED): Point2d -> Point
        # conversion rounds (cvRound, half to even); inclusive corners, clipped to the image.
        h = self.half
        x0, y0 = int(np.rint(x - h)), int(np.rint(y - h))
        x1, y1 = int(np.rint(x + h)), int(np.rint(y + h))
        x0, y0 = max(x0, 0), max(y0, 0)
        x1, y1 = min(x1, self.cols - 1), min(y1, self.rows - 1)
        if x1 >= x0 and y1 >= y0:
            self.m[y0 : y1 + 1, x0 : x1 + 1] = 0


def select_keypoints(mask: Mask, xy, score, num_to_add):
    """Greedy pick of Tracker::DetectLK (tracker.cpp:224-229, :295-328) without the descriptor /
    rescue branches.  Keypoints are first filtered by the mask at (int)(x+0.5),(int)(y+0.5)
    (cv::KeyPointsFilter::runByPixelsMask inside detect(img, kps, mask)), then sorted by
    response.  DEVIATION (documented in DESIGN.md): the reference's std::sort is unstable, so
    its order among equal integer scores is libstdc++-specific; oracle and product both use
    the total order (score desc, y asc, x asc)."""
    keep = [i for i in range(len(xy)) if mask.m[int(xy[i][1] + 0.5), int(xy[i][0] + 0.5)]]
    keep.sort(key=lambda i: (-int(score[i]), int(xy[i][1]), int(xy[i][0])))
    picked = []
    for i in keep:
        x, y = float(xy[i][0]), float(xy[i][1])
        if mask.valid(x, y):
            picked.append(i)
            mask.mask_out(x, y)
            num_to_add -= 1
        if num_to_add <= 0 or score[i] < 5:
            break
    return picked
