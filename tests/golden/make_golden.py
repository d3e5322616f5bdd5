"""Generates tests/golden/*.npz with the cv2 wheel of this image (the only available pin for the
tracker arithmetic).  Run from the repo root:  python tests/golden/make_golden.py
The npz files are committed; the GPU box never needs cv2 or /root/reference to use them."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from xivo_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    cv2.setNumThreads(1)
    rows, cols = 120, 160
    a, b = synth.frame_pair(rows, cols, seed=3, shift=(2, 1))
    a3, b3 = synth.to_bgr(a, True), synth.to_bgr(b, True)
    out = dict(a=a, b=b)
    # pyramid (pyrDown chain, win 15 -> levels until <= 15)
    lv = [a]
    while min((lv[-1].shape[0] + 1) // 2, (lv[-1].shape[1] + 1) // 2) > 15 and len(lv) <= 5:
        lv.append(cv2.pyrDown(lv[-1]))
    for i, l in enumerate(lv):
        out[f"pyr{i}"] = l
    lv3 = [a3]
    for _ in range(len(lv) - 1):
        lv3.append(cv2.pyrDown(lv3[-1]))
    for i, l in enumerate(lv3):
        out[f"pyr3_{i}"] = l
    out["gray_from_bgr"] = cv2.cvtColor(a3, cv2.COLOR_BGR2GRAY)
    # FAST
    for thr in (10, 20):
        kps = cv2.FastFeatureDetector_create(thr, True).detect(a, None)
        out[f"fast{thr}"] = np.array([[k.pt[0], k.pt[1], k.response] for k in kps], np.int32).reshape(-1, 3)
    kps = cv2.FastFeatureDetector_create(20, True).detect(a3, None)
    out["fast20_bgr"] = np.array([[k.pt[0], k.pt[1], k.response] for k in kps], np.int32).reshape(-1, 3)
    # LK
    kps = sorted(cv2.FastFeatureDetector_create(20, True).detect(a, None), key=lambda k: (-k.response, k.pt[1], k.pt[0]))[:60]
    p0 = np.array([k.pt for k in kps], np.float32)
    init = p0 + np.float32([-1.5, -0.5])
    crit = (cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, 30, 0.01)
    for name, (ia, ib) in dict(gray=(a, b), bgr=(a3, b3)).items():
        p1, st, er = cv2.calcOpticalFlowPyrLK(ia, ib, p0, init.copy(), winSize=(15, 15), maxLevel=5, criteria=crit,
                                              flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
        out[f"lk_{name}_p1"], out[f"lk_{name}_st"], out[f"lk_{name}_err"] = p1, st.ravel(), er.ravel()
    out["lk_p0"], out["lk_init"] = p0, init
    np.savez_compressed(os.path.join(OUT, "tracker_cv2.npz"), **out)
    print("wrote tracker_cv2.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
