"""GPU image pipeline with the tracker-level homography outlier rejection (`do_outlier_rejection`, on in the reference's tracker-only
configs; Tracker::OutlierRejection tracker.cpp:594-599, :705-753 -> cv::findHomography).  The mask itself is host work and is pinned on
the CPU (csrc/homography.h against oracle/homography_oracle.py and cv2 4.13: tests/test_host_logic.py, tests/test_capi_symbols.py); here
the accept loop that applies it runs inside the CUDA pipeline.
(File name sorts after the other GPU suites on purpose: this row was added after the last GPU minute of round 1.)"""
import os

import numpy as np
import pytest

from oracle.estimator_oracle import EstimatorOracle
from xivo_b200 import pyxivo, sim

pytestmark = pytest.mark.gpu
CFG = os.path.join(os.path.dirname(pyxivo.__file__), "cfg")


def _cfg(method, thresh):
    cfg = sim.load_cfg(os.path.join(CFG, "vio_640x480.json"))
    cfg["camera_cfg"].update(rows=240, cols=320, fx=137.5, fy=137.5, cx=160, cy=120)
    cfg["tracker_cfg"].update(num_features_min=60, num_features_max=80, do_outlier_rejection=True,
                              outlier_rejection={"method": method, "RANSAC_reproj_thresh": thresh, "RANSAC_max_iters": 2000, "confidence": 0.995})
    return cfg


@pytest.mark.parametrize("method", ["LMEDS", "RANSAC"])
def test_image_pipeline_parity_with_outlier_rejection_enabled(method):
    """The reference's 3 px threshold: on this stream every track stays an inlier, so the pipeline must reproduce the oracle exactly
    (id tables, positions, pose) with the rejection stage in the loop."""
    cfg = _cfg(method, 3.0)
    msgs, _ = sim.image_stream(cfg, duration=1.6, channels=1, seed=1)
    ref = EstimatorOracle(cfg, G=4, F=14)
    b = pyxivo.Batch(cfg, n_seq=1, max_groups=4, max_features=14)
    n = 0
    for kind, ts, p in msgs:
        if kind == "imu":
            ref.InertialMeas(ts, p[0], p[1])
            b.inertial_meas(ts, p[0], p[1])
        else:
            ref.VisualMeas(ts, p)
            b.visual_meas(ts, [p])
            n += 1
            ids, xy, _st = b.tracked_features(0)
            assert ids.tolist() == [f.id for f in ref.tracks], f"frame {n}"
            assert np.abs(b.gsb(0) - ref.gsb()).max() <= 1e-5, f"frame {n}"
            assert b.tracker_counters(0)["num_tracker_outlier_rejected"] == ref.num_outliers_rejected
    assert n > 25 and len(ref.tracks) >= 30
    b.close()


def test_tight_threshold_rejects_tracks_like_the_oracle():
    """A sub-pixel threshold makes the homography reject a few tracks per frame once the camera moves.  Which borderline track goes can
    differ between two LK implementations that agree to 1e-4 px, so only the totals are compared."""
    cfg = _cfg("LMEDS", 0.3)
    msgs, _ = sim.image_stream(cfg, duration=2.4, channels=1, seed=1)
    ref = EstimatorOracle(cfg, G=4, F=14)
    b = pyxivo.Batch(cfg, n_seq=1, max_groups=4, max_features=14)
    tot_p = tot_o = 0
    for kind, ts, p in msgs:
        if kind == "imu":
            ref.InertialMeas(ts, p[0], p[1])
            b.inertial_meas(ts, p[0], p[1])
        else:
            ref.VisualMeas(ts, p)
            b.visual_meas(ts, [p])
            tot_o += ref.num_outliers_rejected
            tot_p += b.tracker_counters(0)["num_tracker_outlier_rejected"]
    # after the first borderline difference the two runs track different feature sets, so the totals only have to be of the same order
    assert tot_o >= 5 and 2 <= tot_p <= 3 * tot_o + 5, (tot_p, tot_o)
    assert len(b.tracked_features(0)[0]) >= 30
    b.close()
