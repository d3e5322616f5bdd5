"""Descriptor path of the tracker: BRIEF-32 per track and frame, descriptor-distance check, rescue of newly
dropped tracks by cross-checked brute-force matching (/root/reference/src/tracker.cpp:231-292, :530-565), and the MATCH tracker
(Tracker::UpdateMatch, :341-460) -- the CUDA pipeline (brief_kernel, hamming_nearest_kernel + host decisions) against the pipeline oracle
(oracle/estimator_oracle.py + oracle/tracker_oracle.c), frame by frame: identical track lists (ids, order), positions, in-state tables,
pose.  The BRIEF test pairs are this repository's own table (opencv_contrib's is not vendored in the reference and cv2.xfeatures2d is
absent): descriptor VALUES are parity-unpinned against OpenCV, the matcher is pinned on cv2.BFMatcher (tests/test_oracle_tracker.py)."""
import os

import numpy as np
import pytest

from test_gpu_estimator import _run_image_parity
from xivo_b200 import pyxivo, sim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = os.path.join(ROOT, "xivo_b200", "cfg")
pytestmark = pytest.mark.gpu


def _cfg(**tracker):
    cfg = sim.load_cfg(os.path.join(CFG, "vio_640x480.json"))
    cfg["camera_cfg"].update(rows=240, cols=320, fx=137.5, fy=137.5, cx=160, cy=120)
    cfg["tracker_cfg"].update(num_features_min=60, num_features_max=80)
    cfg["tracker_cfg"].update(tracker)
    return cfg


@pytest.mark.parametrize("case", ["rescue_like_tumvi_cam0", "distance_check_differential", "rescue_with_outlier_rejection"])
def test_lk_tracker_with_descriptors_matches_the_oracle(case):
    """rescue_like_tumvi_cam0: the descriptor block of the reference's cfg/tumvi_cam0.json:218-239 (extract_descriptor, match_dropped_tracks,
    no distance check, differential off).  distance_check_differential: descriptor_distance_thresh on, descriptors replaced every frame.
    rescue_with_outlier_rejection: both on -- DetectLK's check_homography (which compares pixel positions, tracker.cpp:818-828)."""
    if case == "rescue_like_tumvi_cam0":
        cfg = _cfg(extract_descriptor=True, match_dropped_tracks=True, differential=False, descriptor_distance_thresh=-1, descriptor="BRIEF")
    elif case == "distance_check_differential":
        cfg = _cfg(extract_descriptor=True, match_dropped_tracks=True, differential=True, descriptor_distance_thresh=70)
    else:
        cfg = _cfg(extract_descriptor=True, match_dropped_tracks=True, differential=True, do_outlier_rejection=True,
                   outlier_rejection={"method": "LMEDS", "RANSAC_reproj_thresh": 3.0, "RANSAC_max_iters": 2000, "confidence": 0.995})
    ref, b, nframes = _run_image_parity(cfg, 4, 14, 1.6, 2)
    tc = np.array(ref.track_counts[3:])
    assert nframes >= 35 and tc.max() >= 75 and tc.mean() >= 45
    b.close()


def test_match_tracker_full_vio_matches_the_oracle():
    """tracker_type MATCH (the reference's cfg/phab_tracker_only.json:50 uses it with SIFT; BRIEF here): detection + description of every frame,
    tracks continued by cross-checked nearest-neighbour matching."""
    cfg = _cfg(tracker_type="MATCH", extract_descriptor=True, differential=True, descriptor_distance_thresh=-1)
    ref, b, nframes = _run_image_parity(cfg, 4, 14, 1.6, 2)
    tc = np.array(ref.track_counts[3:])
    assert nframes >= 35 and tc.max() == 80 and tc.mean() >= 50
    b.close()


def test_match_tracker_tracker_only_mode():
    cfg = _cfg(tracker_type="MATCH", extract_descriptor=True, differential=False, descriptor_distance_thresh=90)
    from oracle.estimator_oracle import EstimatorOracle

    msgs, _ = sim.image_stream(cfg, duration=1.0, seed=3)
    ref = EstimatorOracle(cfg, G=4, F=14, tracker_only=True)
    b = pyxivo.Batch(cfg, n_seq=2, max_groups=4, max_features=14, tracker_only=True)
    n = 0
    for kind, ts, p in msgs:
        if kind != "img":
            continue
        ref.VisualMeasTrackerOnly(ts, p)
        b.visual_meas(ts, [p, p], tracker_only=True)
        n += 1
        for s in range(2):
            ids, xy, _ = b.tracked_features(s)
            assert ids.tolist() == [f.id for f in ref.tracks], f"frame {n} sequence {s}"
            if len(ids):
                assert np.abs(xy - np.array([f.xp() for f in ref.tracks])).max() <= 1e-3
    assert n >= 20 and len(ref.tracks) >= 50
    b.close()
