"""Device-side tracker decisions (track_accept_kernel / track_select_kernel, xivo_b200/csrc/tracker_kernels.cu) against the host
implementation of the same loops (Tracker::UpdateLK's accept loop and Tracker::DetectLK's greedy selection, reference
src/tracker.cpp:571-589, :295-328), which the pipeline oracle pins: XIVO_HOST_TRACKER_DECISIONS=1 routes a batch through the host
code, the default through the kernels; track lists (ids, order, positions) must be identical frame by frame.  The inputs are chosen to
stress what the pipeline suites rarely reach: thousands of keypoints with few distinct scores (ties resolved by (y, x); several
1024-key chunks of the radix select), budgets larger than one chunk, margins and mask blocks clipped at the image border, a frame on
which nothing can be tracked (re-initialisation), and 3-channel input."""
import os

import numpy as np
import pytest

from xivo_b200 import pyxivo, sim, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = os.path.join(ROOT, "xivo_b200", "cfg")
pytestmark = pytest.mark.gpu


def _cfg(rows, cols, nmin, nmax, thr, mask=15, margin=8):
    cfg = sim.load_cfg(os.path.join(CFG, "vio_640x480.json"))
    cfg["camera_cfg"].update(rows=rows, cols=cols, fx=cols * 0.43, fy=cols * 0.43, cx=cols / 2, cy=rows / 2)
    cfg["tracker_cfg"].update(num_features_min=nmin, num_features_max=nmax, mask_size=mask, margin=margin)
    cfg["tracker_cfg"]["FAST"] = {"threshold": thr, "nonmaxSuppression": True}
    cfg["message_buffer_size"] = 0
    return cfg


def _run(cfg, frames, host_decisions, n_seq=3):
    os.environ["XIVO_HOST_TRACKER_DECISIONS"] = "1" if host_decisions else "0"
    try:
        b = pyxivo.Batch(cfg, n_seq=n_seq, max_groups=4, max_features=14, tracker_only=True)
        out = []
        for k, img in enumerate(frames):
            b.visual_meas(k * 40_000_000, [img] * n_seq, tracker_only=True)
            out.append([b.tracked_features(s) for s in range(n_seq)])
        b.close()
    finally:
        os.environ.pop("XIVO_HOST_TRACKER_DECISIONS", None)
    return out


def _same(a, b):
    assert len(a) == len(b)
    n_tracks = []
    for k, (fa, fb) in enumerate(zip(a, b)):
        for s, ((ia, xa, sa), (ib, xb, sb)) in enumerate(zip(fa, fb)):
            assert ia.tolist() == ib.tolist(), f"frame {k} sequence {s}: ids differ"
            assert np.array_equal(xa, xb) and np.array_equal(sa, sb), f"frame {k} sequence {s}"
        n_tracks.append(len(fa[0][0]))
    return n_tracks


@pytest.mark.parametrize("case", ["noise_many_ties", "texture_big_budget", "small_mask_wide_margin", "bgr"])
def test_device_decisions_equal_host_decisions(case):
    rng = np.random.default_rng(7)
    if case == "noise_many_ties":  # uniform noise, low threshold: > 5000 keypoints, scores in a narrow band -> many (y, x) tie-breaks
        cfg = _cfg(240, 320, 500, 700, 5, mask=5, margin=4)
        base = rng.integers(90, 166, (260, 340), dtype=np.uint8)
        frames = [base[k : k + 240, 2 * k : 2 * k + 320].copy() for k in range(6)]
    elif case == "texture_big_budget":  # budget 1500 > one 1024-key chunk of the select kernel
        cfg = _cfg(480, 640, 1200, 1500, 10, mask=7, margin=8)
        canvas = synth.texture_canvas(480, 640, seed=2, pad=64)
        frames = [synth.frame_from_canvas(canvas, 480, 640, (2 * k, k), noise_seed=20 + k, pad=32) for k in range(6)]
    elif case == "small_mask_wide_margin":
        cfg = _cfg(240, 320, 60, 90, 20, mask=31, margin=40)
        canvas = synth.texture_canvas(240, 320, seed=3, pad=64)
        frames = [synth.frame_from_canvas(canvas, 240, 320, (3 * k, 2 * k), noise_seed=30 + k, pad=32) for k in range(5)]
        frames.insert(3, np.full((240, 320), 128, np.uint8))  # a blank frame: every track is lost, the next frame re-initialises
    else:
        cfg = _cfg(240, 320, 75, 100, 20)
        canvas = synth.texture_canvas(240, 320, seed=4, pad=64)
        frames = [synth.to_bgr(synth.frame_from_canvas(canvas, 240, 320, (2 * k, 2 * k), noise_seed=40 + k, pad=32), distinct=True) for k in range(6)]
    dev = _run(cfg, frames, host_decisions=False)
    host = _run(cfg, frames, host_decisions=True)
    n = _same(dev, host)
    print(case, "tracks per frame", n)
    assert max(n) > 50
    if case == "texture_big_budget":
        assert n[0] > 1024  # more picks than one chunk holds


@pytest.mark.parametrize("size", [(240, 320), (480, 640), (176, 208)])
def test_tma_pyramid_equals_thread_staged_pyramid(size):
    """pyrdown_tma_kernel (cp.async.bulk.tensor box loads, REFLECT_101 patched on the rim CTAs) against pyrdown_vec_kernel: the tracks
    of a short sequence — LK reads every pyramid level, also next to the image border (margin 4) — must be bit-identical.
    (176, 208): a width whose coarser levels stop being multiples of 16, so TMA and thread-staged passes mix inside one pyramid."""
    rows, cols = size
    cfg = _cfg(rows, cols, 150, 200, 10, mask=9, margin=4)
    canvas = synth.texture_canvas(rows, cols, seed=5, pad=64)
    frames = [synth.frame_from_canvas(canvas, rows, cols, (3 * k, 2 * k), noise_seed=50 + k, pad=32) for k in range(6)]

    def run(tma):
        os.environ["XIVO_PYRDOWN_TMA"] = os.environ["XIVO_FAST_TMA"] = "1" if tma else "0"  # both TMA passes (pyramid, FAST tiles) on / off
        try:
            return _run(cfg, frames, host_decisions=False, n_seq=2)
        finally:
            os.environ.pop("XIVO_PYRDOWN_TMA", None)
            os.environ.pop("XIVO_FAST_TMA", None)

    n = _same(run(True), run(False))
    assert max(n) > 100
