"""pyxivo-compatible Python surface over the estimator-level C ABI (include/xivo_b200_estimator.h).

`Estimator` mirrors the method names of the reference's pybind11 module
(/root/reference/pybind11/pyxivo.cpp:332-398) for the hot path; `Batch` is the lock-step
multi-sequence form that the B200 needs to be kept busy.  ctypes plumbing only — all computation
happens in libxivo_b200.so.
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

from . import capi
from .capi import XivoError, _check, _p
from .sim import strip_json_comments


def _resolve_cfg(cfg, base_dir=None) -> str:
    """Accepts a path, a JSON(-with-comments) string or a dict; inlines camera_cfg / tracker_cfg
    given as paths (src/factory.cpp:31-45) and returns strict JSON text."""
    if isinstance(cfg, dict):
        d = dict(cfg)
    else:
        if "{" not in cfg:
            base_dir = base_dir or os.path.dirname(os.path.abspath(cfg))
            cfg = open(cfg).read()
        d = json.loads(strip_json_comments(cfg))
    for k in ("camera_cfg", "tracker_cfg"):
        if isinstance(d.get(k), str):
            path = d[k] if os.path.isabs(d[k]) or base_dir is None else os.path.join(base_dir, d[k])
            if not os.path.exists(path) and base_dir:
                path = os.path.join(os.path.dirname(base_dir), d[k])
            d[k] = json.loads(strip_json_comments(open(path).read()))
    return json.dumps(d)


class Batch:
    def __init__(self, cfg, n_seq=1, max_groups=15, max_features=30, tracker_only=False, device=0, ctx=None, overrides=None):
        text = _resolve_cfg(cfg)
        if overrides:
            d = json.loads(text)
            for k, v in overrides.items():
                if isinstance(v, dict) and isinstance(d.get(k), dict):
                    d[k].update(v)
                else:
                    d[k] = v
            text = json.dumps(d)
        self.cfg = json.loads(text)
        self._own_ctx = ctx is None
        self.ctx = ctx or capi.Context(device)
        self._h = C.c_void_p()
        L = capi.lib()
        _check(L.xivo_batch_create(self.ctx._h, text.encode(), int(n_seq), int(max_groups), int(max_features), int(bool(tracker_only)),
                                   C.byref(self._h)), "xivo_batch_create")
        self.n = n_seq
        self.G, self.F = max_groups, max_features
        self.N = L.xivo_batch_state_dim(self._h)
        self._keep = None

    def close(self):
        if self._h:
            capi.lib().xivo_batch_destroy(self._h)
            self._h = C.c_void_p()
        if self._own_ctx and self.ctx:
            self.ctx.close()
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- ingest -------------------------------------------------------------------------
    def inertial_meas(self, ts_ns, gyro, accel):
        ts = np.ascontiguousarray(np.broadcast_to(np.asarray(ts_ns, dtype=np.uint64), (self.n,)))
        g = np.ascontiguousarray(np.broadcast_to(np.asarray(gyro, dtype=np.float64), (self.n, 3)))
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(accel, dtype=np.float64), (self.n, 3)))
        _check(capi.lib().xivo_batch_inertial_meas(self._h, _p(ts), _p(g), _p(a)), "xivo_batch_inertial_meas")

    def visual_meas(self, ts_ns, imgs, tracker_only=False):
        """imgs: list of n arrays (rows x cols [x 3], uint8, C-contiguous) or one array shared by all."""
        if isinstance(imgs, np.ndarray):
            imgs = [imgs] * self.n
        imgs = [np.ascontiguousarray(i, dtype=np.uint8) for i in imgs]
        rows, cols = imgs[0].shape[:2]
        ch = 1 if imgs[0].ndim == 2 else imgs[0].shape[2]
        ts = np.ascontiguousarray(np.broadcast_to(np.asarray(ts_ns, dtype=np.uint64), (self.n,)))
        ptrs = (C.c_void_p * self.n)(*[i.ctypes.data for i in imgs])
        self._keep = imgs
        _check(capi.lib().xivo_batch_visual_meas(self._h, _p(ts), ptrs, rows, cols, ch, int(bool(tracker_only))), "xivo_batch_visual_meas")

    def step(self, imu_ts, gyro, accel, frame_ts, imgs):
        """n_imu InertialMeas + one VisualMeas per sequence in one call (xivo_batch_step).
        imu_ts: (n_imu, n) or (n_imu,), gyro/accel: (n_imu, n, 3) or (n_imu, 3)."""
        if isinstance(imgs, np.ndarray):
            imgs = [imgs] * self.n
        imgs = [np.ascontiguousarray(i, dtype=np.uint8) for i in imgs]
        rows, cols = imgs[0].shape[:2]
        ch = 1 if imgs[0].ndim == 2 else imgs[0].shape[2]
        its = np.asarray(imu_ts, dtype=np.uint64)
        n_imu = its.shape[0]
        its = np.ascontiguousarray(np.broadcast_to(its.reshape(n_imu, -1), (n_imu, self.n)))
        g = np.ascontiguousarray(np.broadcast_to(np.asarray(gyro, dtype=np.float64).reshape(n_imu, -1, 3), (n_imu, self.n, 3)))
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(accel, dtype=np.float64).reshape(n_imu, -1, 3), (n_imu, self.n, 3)))
        fts = np.ascontiguousarray(np.broadcast_to(np.asarray(frame_ts, dtype=np.uint64), (self.n,)))
        ptrs = (C.c_void_p * self.n)(*[i.ctypes.data for i in imgs])
        self._keep = imgs
        _check(capi.lib().xivo_batch_step(self._h, n_imu, _p(its), _p(g), _p(a), _p(fts), ptrs, rows, cols, ch, 0), "xivo_batch_step")

    def visual_meas_pointcloud(self, ts_ns, ids, xp_depth, tracker_only=False):
        """ids / xp_depth: lists of per-sequence arrays, or single arrays shared by all."""
        if isinstance(ids, np.ndarray):
            ids, xp_depth = [ids] * self.n, [xp_depth] * self.n
        ids = [np.ascontiguousarray(i, dtype=np.int32) for i in ids]
        xpd = [np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 3) for x in xp_depth]
        npts = np.array([len(i) for i in ids], dtype=np.int32)
        ts = np.ascontiguousarray(np.broadcast_to(np.asarray(ts_ns, dtype=np.uint64), (self.n,)))
        pi = (C.c_void_p * self.n)(*[i.ctypes.data for i in ids])
        px = (C.c_void_p * self.n)(*[x.ctypes.data for x in xpd])
        _check(capi.lib().xivo_batch_visual_meas_pointcloud(self._h, _p(ts), _p(npts), pi, px, int(bool(tracker_only))),
               "xivo_batch_visual_meas_pointcloud")

    def init_with_sim_depths(self):
        _check(capi.lib().xivo_init_with_sim_depths(self._h), "xivo_init_with_sim_depths")

    # ---- read-back ----------------------------------------------------------------------
    def _g(self, fn, seq):
        out = np.zeros(12)
        _check(getattr(capi.lib(), fn)(self._h, seq, _p(out)), fn)
        return out.reshape(3, 4)

    def gsb(self, seq=0):
        return self._g("xivo_get_gsb", seq)

    def gbc(self, seq=0):
        return self._g("xivo_get_gbc", seq)

    def gsc(self, seq=0):
        return self._g("xivo_get_gsc", seq)

    def motion(self, seq=0):
        V, bg, ba, Rsg = np.zeros(3), np.zeros(3), np.zeros(3), np.zeros(9)
        _check(capi.lib().xivo_get_motion(self._h, seq, _p(V), _p(bg), _p(ba), _p(Rsg)), "xivo_get_motion")
        return V, bg, ba, Rsg.reshape(3, 3)

    def P(self, seq=0):
        out = np.zeros((self.N, self.N))
        _check(capi.lib().xivo_get_P(self._h, seq, _p(out)), "xivo_get_P")
        return out

    def Pstate(self, seq=0):
        out = np.zeros((9, 9))
        _check(capi.lib().xivo_get_Pstate(self._h, seq, _p(out)), "xivo_get_Pstate")
        return out

    COUNTER_NAMES = ("num_instate_features", "num_instate_groups", "gauge_group", "num_mh_rejected", "num_tracker_failed",
                     "num_tracker_new_detections", "vision_counter", "imu_counter", "MeasurementUpdateInitialized",
                     "VisionInitialized", "num_tracked", "error")

    def counters(self, seq=0):
        out = np.zeros(12, np.int32)
        _check(capi.lib().xivo_get_counters(self._h, seq, _p(out)), "xivo_get_counters")
        return dict(zip(self.COUNTER_NAMES, out.tolist()))

    def now(self, seq=0):
        ts = C.c_uint64()
        _check(capi.lib().xivo_get_time_ns(self._h, seq, C.byref(ts)), "xivo_get_time_ns")
        return ts.value

    def tracked_features(self, seq=0, max_n=4096):
        ids, xy, st, n = np.zeros(max_n, np.int32), np.zeros((max_n, 2)), np.zeros(max_n, np.int32), C.c_int()
        _check(capi.lib().xivo_get_tracked_features(self._h, seq, _p(ids), _p(xy), _p(st), max_n, C.byref(n)), "xivo_get_tracked_features")
        k = min(n.value, max_n)
        return ids[:k].copy(), xy[:k].copy(), st[:k].copy()

    def instate_features(self, seq=0):
        m = self.F
        ids, sinds, refs, Xs, x, n = np.zeros(m, np.int32), np.zeros(m, np.int32), np.zeros(m, np.int32), np.zeros((m, 3)), np.zeros((m, 3)), C.c_int()
        _check(capi.lib().xivo_get_instate_features(self._h, seq, _p(ids), _p(sinds), _p(refs), _p(Xs), _p(x), m, C.byref(n)),
               "xivo_get_instate_features")
        k = min(n.value, m)
        return dict(ids=ids[:k].copy(), sinds=sinds[:k].copy(), ref_groups=refs[:k].copy(), Xs=Xs[:k].copy(), x=x[:k].copy())

    def instate_groups(self, seq=0):
        m = self.G
        ids, sinds, g, n = np.zeros(m, np.int32), np.zeros(m, np.int32), np.zeros((m, 12)), C.c_int()
        _check(capi.lib().xivo_get_instate_groups(self._h, seq, _p(ids), _p(sinds), _p(g), m, C.byref(n)), "xivo_get_instate_groups")
        k = min(n.value, m)
        return dict(ids=ids[:k].copy(), sinds=sinds[:k].copy(), gsb=g[:k].reshape(k, 3, 4).copy())


class Estimator:
    """Single-sequence facade with the reference's pyxivo method names
    (pybind11/pyxivo.cpp:332-398): pyxivo.Estimator(cfg_path, viewer_cfg_path, name, tracker_only)."""

    def __init__(self, cfg, viewer_cfg="", name="", tracker_only=False, max_groups=15, max_features=30, device=0):
        self._b = Batch(cfg, 1, max_groups, max_features, tracker_only, device)
        self._tracker_only = tracker_only

    def InertialMeas(self, ts, wx, wy, wz, ax, ay, az):
        self._b.inertial_meas(int(ts), [wx, wy, wz], [ax, ay, az])

    def VisualMeas(self, ts, image):
        if isinstance(image, str):
            raise XivoError("VisualMeas(path) needs an image decoder; pass the decoded uint8 array instead")
        self._b.visual_meas(int(ts), [image], tracker_only=False)

    def VisualMeasTrackerOnly(self, ts, image):
        self._b.visual_meas(int(ts), [image], tracker_only=True)

    def VisualMeasPointCloud(self, ts, feature_ids, xp_and_depths):
        self._b.visual_meas_pointcloud(int(ts), [np.asarray(feature_ids)], [np.asarray(xp_and_depths)], tracker_only=False)

    def VisualMeasPointCloudTrackerOnly(self, ts, feature_ids, xp_and_depths):
        self._b.visual_meas_pointcloud(int(ts), [np.asarray(feature_ids)], [np.asarray(xp_and_depths)], tracker_only=True)

    def InitWithSimDepths(self):
        self._b.init_with_sim_depths()

    def CloseLoop(self):
        return None  # USE_MAPPER is off in the default build (src/CMakeLists.txt:18)

    def gsb(self):
        return self._b.gsb()

    def gsc(self):
        return self._b.gsc()

    def gbc(self):
        return self._b.gbc()

    def Vsb(self):
        return self._b.motion()[0]

    def bg(self):
        return self._b.motion()[1]

    def ba(self):
        return self._b.motion()[2]

    def Rsg(self):
        return self._b.motion()[3]

    def P(self):
        return self._b.P()

    def Pstate(self):
        return self._b.Pstate()

    def now(self):
        return self._b.now()

    def ts(self):
        return self._b.now()

    def gauge_group(self):
        return self._b.counters()["gauge_group"]

    def num_instate_features(self):
        return self._b.counters()["num_instate_features"]

    def num_instate_groups(self):
        return self._b.counters()["num_instate_groups"]

    def num_mh_rejected(self):
        return self._b.counters()["num_mh_rejected"]

    def num_tracker_failed(self):
        return self._b.counters()["num_tracker_failed"]

    def num_tracker_new_detections(self):
        return self._b.counters()["num_tracker_new_detections"]

    def MeasurementUpdateInitialized(self):
        return bool(self._b.counters()["MeasurementUpdateInitialized"])

    def VisionInitialized(self):
        return bool(self._b.counters()["VisionInitialized"])

    def UsingLoopClosure(self):
        return False

    def InstateFeatureIDs(self):
        return self._b.instate_features()["ids"]

    def InstateFeatureSinds(self):
        return self._b.instate_features()["sinds"]

    def InstateFeaturePositions(self):
        return self._b.instate_features()["Xs"]

    def InstateGroupIDs(self):
        return self._b.instate_groups()["ids"]

    def InstateGroupSinds(self):
        return self._b.instate_groups()["sinds"]

    def InstateGroupPoses(self):
        return self._b.instate_groups()["gsb"]

    def tracked_features_no_descriptor(self):
        ids, xy, _ = self._b.tracked_features()
        return [(int(i), p) for i, p in zip(ids, xy)]

    def close(self):
        self._b.close()
