"""oracle/cpu_baseline.py — TEST / BASELINE INFRASTRUCTURE ONLY.

Times the reference's CPU arithmetic for the hot path on the host cores: the
restated pipeline (oracle/estimator_oracle.py) is run to obtain, for every frame, exactly the
inputs the reference would hand to its third-party numerics, and ONLY those numerics are timed with
the fastest faithful CPU implementation available:
  * tracker  : cv2.calcOpticalFlowPyrLK / cv2.FastFeatureDetector (the OpenCV calls of
               src/tracker.cpp:224, :526; cv2.setNumThreads(1), parallelism comes from processes),
  * EKF      : oracle/_ref/libekf_eigen.so = the reference's Eigen 3.3.9 expression sequence for
               MHGating + UpdateJosephForm (single-threaded, like the reference); numpy if absent.
Python bookkeeping, propagation, sub-filter and Jacobian assembly are NOT charged to that figure
(`fps`, kind "port": an upper bound on the reference's frames/s).

Since the reference's own estimator builds here (oracle/build_ref.py -> oracle/_ref/libxivo_ref_G<g>_F<f>.so, its unmodified
sources), every worker additionally runs THAT library on a point-cloud stream with the same state size and about the same
number of tracked features and measures its wall time per frame (8 IMU messages + 1 visual message: propagation, ProcessTracks,
sub-filters, Jacobians, gating, update, feature/group management — everything but the image tracker, which needs OpenCV C++).
`fps_reference` = frames/s with a frame costing (cv2 tracker calls on the image stream) + (reference estimator on the point-cloud
stream): kind "reference" for the estimator half, OpenCV's own code (through cv2) for the tracker half."""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class StageTimer:
    def __init__(self):
        self.frame_s = 0.0
        self.stage = dict(lk=0.0, fast=0.0, gate=0.0, update=0.0)
        try:
            import cv2

            cv2.setNumThreads(1)
            self.cv2 = cv2
        except Exception:  # pragma: no cover
            self.cv2 = None
        so = os.path.join(_HERE, "_ref", "libekf_eigen.so")
        self.eig = None
        if os.path.exists(so):
            self.eig = C.CDLL(so)
            self.eig.ref_update_joseph.restype = C.c_double
            self.eig.ref_mh_gating.restype = C.c_double
        self.kind = "port"

    def _add(self, k, dt):
        self.stage[k] += dt
        self.frame_s += dt

    def lk(self, prev, img, p0, p1, klt):
        if self.cv2 is None:
            return
        cv2 = self.cv2
        crit = (cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, klt["max_iter"], klt["eps"])
        t = time.perf_counter()
        cv2.calcOpticalFlowPyrLK(prev, img, p0, p1.copy(), winSize=(klt["win"], klt["win"]), maxLevel=klt["max_level"], criteria=crit,
                                 flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
        self._add("lk", time.perf_counter() - t)

    def fast(self, img, thr, nms):
        if self.cv2 is None:
            return
        det = self.cv2.FastFeatureDetector_create(int(thr), bool(nms))
        t = time.perf_counter()
        det.detect(img, None)
        self._add("fast", time.perf_counter() - t)

    def gate(self, J, P, inn, R):
        n, _, N = J.shape
        if self.eig is not None:
            J, P, inn = np.ascontiguousarray(J), np.ascontiguousarray(P), np.ascontiguousarray(inn)
            d = np.zeros(n)
            vp = lambda a: a.ctypes.data_as(C.c_void_p)
            self._add("gate", self.eig.ref_mh_gating(N, n, vp(J), vp(P), vp(inn), C.c_double(R), vp(d)))
        else:
            t = time.perf_counter()
            for i in range(n):
                S = J[i] @ P @ J[i].T + R * np.eye(2)
                inn[i] @ np.linalg.solve(S, inn[i])
            self._add("gate", time.perf_counter() - t)

    def update(self, H, P, inn, diagR):
        M, N = H.shape
        if self.eig is not None:
            H, P2, inn, diagR = np.ascontiguousarray(H), np.array(P, order="C", copy=True), np.ascontiguousarray(inn), np.ascontiguousarray(diagR)
            err = np.zeros(N)
            vp = lambda a: a.ctypes.data_as(C.c_void_p)
            self._add("update", self.eig.ref_update_joseph(N, M, vp(H), vp(P2), vp(inn), vp(diagR), vp(err)))
        else:
            from . import ekf_oracle as E

            t = time.perf_counter()
            E.update_joseph(H, P, inn, diagR)
            self._add("update", time.perf_counter() - t)


def worker(args):
    """Runs one synthetic sequence through the restated pipeline and returns the per-frame time of the
    timed numerics for the frames after `skip` executed frames."""
    cfg, seed, n_frames, skip, G, F, channels = args
    from xivo_b200 import sim

    from .estimator_oracle import EstimatorOracle

    duration = (n_frames + 12) * 0.04
    msgs, _ = sim.image_stream(cfg, duration=duration, seed=seed, channels=channels, fast=True)
    est = EstimatorOracle(cfg, G=G, F=F)
    st = StageTimer()
    est.stage_timer = st
    per_frame, executed = [], 0
    last = 0.0
    for kind, ts, p in msgs:
        before = est.vision_counter
        if kind == "imu":
            est.InertialMeas(ts, p[0], p[1])
        else:
            est.VisualMeas(ts, p)
        if est.vision_counter != before:  # a frame left the message heap and was executed
            executed += 1
            if executed > skip:
                per_frame.append(st.frame_s - last)
            last = st.frame_s
        if len(per_frame) >= n_frames:
            break
    ref_ms, ref_tracks = None, None
    try:
        ref_ms, ref_tracks = reference_estimator_ms(G, F, n_frames, skip, seed)
    except Exception:  # the library is optional: without it only the numerics-only figure is reported
        ref_ms = None
    return dict(per_frame=per_frame, stage=st.stage, eigen=st.eig is not None, cv2=st.cv2 is not None,
                ninstate=len(est.instate_features), ntracks=len(est.tracks), ref_ms=ref_ms, ref_tracks=ref_tracks)


def reference_estimator_ms(G, F, n_frames, skip, seed):
    """Wall time per frame [ms] of the reference's own estimator (oracle/_ref) on a point-cloud stream: (mean over the frames after
    `skip`, tracked features per frame).  One estimator per process (the reference's singletons): call once per worker process."""
    from xivo_b200 import sim

    from . import ref_runner

    if not ref_runner.available(G, F):
        return None, None
    cfg = sim.load_cfg(os.path.join(os.path.dirname(_HERE), "xivo_b200", "cfg", "pcw_sim.json"))
    cfg["tracker_cfg"].update(num_features_min=120, num_features_max=150)  # Tracker::UpdatePointCloud keeps at most 150 tracks, like the image workload
    msgs, _ = sim.pcw_stream(cfg, duration=(n_frames + skip + 12) * 0.04, seed=seed)
    marks = []
    r = ref_runner.run(cfg, msgs, G, F, True, lib_file=ref_runner.lib_path(G, F, timing=True), on_visual=lambda: marks.append(time.perf_counter()))
    k0 = min(skip + 10, len(marks) - 2)  # +10: the reorder buffer holds the first messages back
    k1 = min(k0 + n_frames, len(marks) - 1)
    per_frame_ms = 1e3 * (marks[k1] - marks[k0]) / max(1, k1 - k0)
    ntr = float(np.mean([min(150, len(p[0])) for kind, _, p in msgs if kind == "pc"]))
    assert r["n_instate"][-1] > 0
    return per_frame_ms, ntr


def run(cfg, n_procs, n_frames, skip, G, F, channels=1):
    """All `n_procs` sequences run concurrently (one process each).  Returns aggregate frames/s of the
    timed numerics and a description."""
    import multiprocessing as mp

    ctx = mp.get_context("fork")
    with ctx.Pool(n_procs, maxtasksperchild=1) as pool:  # one task per process: the reference library keeps process-wide singletons
        res = pool.map(worker, [(cfg, s, n_frames, skip, G, F, channels) for s in range(n_procs)], chunksize=1)
    fps = sum(len(r["per_frame"]) / max(sum(r["per_frame"]), 1e-12) for r in res)
    mean_ms = 1e3 * float(np.mean([np.mean(r["per_frame"]) for r in res]))
    stage = {k: float(np.mean([r["stage"][k] for r in res])) for k in res[0]["stage"]}
    tot = sum(stage.values()) or 1.0
    out = dict(fps=fps, mean_frame_ms=mean_ms, stage_share={k: v / tot for k, v in stage.items()}, eigen=res[0]["eigen"], cv2=res[0]["cv2"],
               frames=sum(len(r["per_frame"]) for r in res), ninstate=res[0]["ninstate"], ntracks=res[0]["ntracks"], fps_reference=None)
    if all(r.get("ref_ms") for r in res):
        # per worker: a frame = the cv2 tracker calls measured on its image stream + the reference's own estimator measured on its point-cloud stream
        fr = [len(r["per_frame"]) for r in res]
        trk_ms = [1e3 * (r["stage"]["lk"] + r["stage"]["fast"]) / max(1, (len(r["per_frame"]) + skip)) for r in res]
        out["fps_reference"] = float(sum(1e3 / (t + r["ref_ms"]) for t, r in zip(trk_ms, res)))
        out["ref_estimator_ms"] = float(np.mean([r["ref_ms"] for r in res]))
        out["tracker_ms"] = float(np.mean(trk_ms))
        out["ref_tracks"] = float(np.mean([r["ref_tracks"] for r in res]))
        del fr
    return out


if __name__ == "__main__":  # python -m oracle.cpu_baseline <cfg.json> <procs> <frames> <skip> <G> <F>  -> JSON on stdout
    import json
    import sys

    cfg_path, procs, frames, skip, G_, F_ = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
    channels_ = int(sys.argv[7]) if len(sys.argv) > 7 else 1
    sys.path.insert(0, os.path.dirname(_HERE))
    from xivo_b200 import sim

    print(json.dumps(run(sim.load_cfg(cfg_path), procs, frames, skip, G_, F_, channels_)))
