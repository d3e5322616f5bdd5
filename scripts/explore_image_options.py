"""Round-2 exploration (GPU box): the image path with the options that were widened on the CPU at the end of round 1 —
triangulate_pre_subfilter, use_depth_opt (two_view), do_outlier_rejection — product vs oracle, frame by frame, printing the first
divergence instead of asserting.  On the image path the CUDA LK and the C oracle agree to 1e-4 px, which a threshold decision
(triangulation window / angular checks, homography mask) can turn into a different feature set; this script measures how often that
happens before such a case becomes a parity test.   usage: python scripts/explore_image_options.py > gpurun_out/r02_explore.txt"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.estimator_oracle import EstimatorOracle  # noqa: E402
from xivo_b200 import sim  # noqa: E402

CFG = os.path.join(ROOT, "xivo_b200", "cfg")
TRI = {"triangulate_pre_subfilter": True, "initial_std_x_badtri": 1.0, "initial_std_y_badtri": 1.0, "initial_std_z_badtri": 1.0,
       "triangulation": {"method": "l1_angular", "zmin": 0.05, "zmax": 5.0, "max_theta_thresh": 0.1, "beta_thesh": 0.25}}
DOPT = {"use_depth_opt": True, "depth_opt": {"two_view": True, "use_hessian": True, "max_iters": 5, "eps": 1e-3, "damping": 1e-3, "max_res_norm": 2.5}}


def make_cfg(over, tracker_over=None):
    cfg = sim.load_cfg(os.path.join(CFG, "vio_640x480.json"))
    cfg["camera_cfg"].update(rows=240, cols=320, fx=137.5, fy=137.5, cx=160, cy=120)
    cfg["tracker_cfg"].update(num_features_min=60, num_features_max=80, **(tracker_over or {}))
    cfg.update(over)
    return cfg


def run(name, over, tracker_over=None, duration=2.0, seed=1):
    from xivo_b200 import pyxivo

    cfg = make_cfg(over, tracker_over)
    msgs, _ = sim.image_stream(cfg, duration=duration, channels=1, seed=seed)
    ref = EstimatorOracle(cfg, G=4, F=14)
    b = pyxivo.Batch(cfg, n_seq=1, max_groups=4, max_features=14)
    n, first, dpose = 0, None, 0.0
    for kind, ts, p in msgs:
        if kind == "imu":
            ref.InertialMeas(ts, p[0], p[1])
            b.inertial_meas(ts, p[0], p[1])
            continue
        ref.VisualMeas(ts, p)
        b.visual_meas(ts, [p])
        n += 1
        ids = b.tracked_features(0)[0].tolist()
        same_tracks = ids == [f.id for f in ref.tracks]
        same_state = sorted(b.instate_features(0)["ids"].tolist()) == sorted(f.id for f in ref.instate_features)
        dpose = float(np.abs(b.gsb(0) - ref.gsb()).max())
        if first is None and not (same_tracks and same_state and dpose <= 1e-5):
            first = (n, same_tracks, same_state, dpose)
    print(f"{name}: {n} frames, first divergence: {first}, final |dpose| {dpose:.2e}, good/bad triangulations {ref.num_good_tri}/{ref.num_bad_tri}, "
          f"refined {ref.num_refined}/{ref.num_refine_failed}, outliers rejected (last frame) {ref.num_outliers_rejected}", flush=True)
    b.close()


if __name__ == "__main__":
    run("baseline", {})
    run("triangulation l1", TRI)
    run("triangulation l2", dict(TRI, triangulation=dict(TRI["triangulation"], method="l2_angular")))
    run("depth refinement (two views)", DOPT)
    run("triangulation + depth refinement", dict(TRI, **DOPT))
    run("LMEDS outlier rejection 3 px", {}, dict(do_outlier_rejection=True, outlier_rejection={"method": "LMEDS", "RANSAC_reproj_thresh": 3.0}))
    run("LMEDS outlier rejection 0.5 px", {}, dict(do_outlier_rejection=True, outlier_rejection={"method": "LMEDS", "RANSAC_reproj_thresh": 0.5}))
