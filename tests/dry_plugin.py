"""pytest plugin for a CPU DRY RUN of the GPU test bodies (authoring aid, not a parity check and never loaded by default):

    PYTHONPATH=tests python -m pytest -p dry_plugin tests/test_gpu_widen_2_triangulation.py ...

replaces xivo_b200.pyxivo.Batch by a stand-in backed by the numpy pipeline oracle, so that the test CODE — keys, shapes, call
signatures, fixtures, golden lookups — executes without a GPU.  What passes here says nothing about the product (the stand-in IS the
oracle); it only catches mistakes in GPU tests written while no GPU is available.  Known stand-in gaps: the cached Feature::Xs_ and
JustDroppedFeatureIDs are not modelled, errors are the oracle's exceptions rather than XivoError."""
import math, os, sys
import numpy as np
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)
from oracle.estimator_oracle import EstimatorOracle, F_GAUGE
from oracle import ekf_oracle as E
from xivo_b200 import pyxivo


def quat(R):
    m = R; t = m[0, 0] + m[1, 1] + m[2, 2]; q = np.zeros(4)
    if t > 0:
        t = math.sqrt(t + 1.0); q[3] = 0.5 * t; t = 0.5 / t
        q[0] = (m[2, 1] - m[1, 2]) * t; q[1] = (m[0, 2] - m[2, 0]) * t; q[2] = (m[1, 0] - m[0, 1]) * t
    else:
        i = 0
        if m[1, 1] > m[0, 0]: i = 1
        if m[2, 2] > m[i, i]: i = 2
        j = (i + 1) % 3; k = (j + 1) % 3
        t = math.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0); q[i] = 0.5 * t; t = 0.5 / t
        q[3] = (m[k, j] - m[j, k]) * t; q[j] = (m[j, i] + m[i, j]) * t; q[k] = (m[k, i] + m[i, k]) * t
    return q


class FakeBatch:
    def __init__(self, cfg, n_seq=1, max_groups=15, max_features=30, tracker_only=False, device=0, overrides=None, ctx=None):
        cfg = dict(cfg)
        if overrides: cfg.update(overrides)
        if "imu_calib" not in cfg:  # reference-shaped tracker-only config (camera + tracker block): the oracle wants the filter sections too
            from xivo_b200 import sim
            full = sim.load_cfg(os.path.join(_ROOT, "xivo_b200", "cfg", "vio_640x480.json"))
            full["tracker_cfg"] = cfg["tracker_cfg"]
            full["camera_cfg"] = dict(dict(fx=190.98, fy=190.97, cx=254.93, cy=256.90), **cfg["camera_cfg"])
            full.update({k: v for k, v in cfg.items() if k not in ("camera_cfg", "tracker_cfg")})
            cfg = full
        tr = cfg.get("triangulation", {})
        if cfg.get("triangulate_pre_subfilter") and tr.get("method", "l1_angular") not in E.TRI_METHODS:
            raise pyxivo.XivoError("batch_create: Incorrect Method for Triangulation: " + tr["method"])
        self.n, self.G, self.F = n_seq, max_groups, max_features
        self.N = 23 + 6 * max_groups + 3 * max_features
        self.e = [EstimatorOracle(cfg, G=max_groups, F=max_features, tracker_only=tracker_only) for _ in range(n_seq)]
        self.just = [[] for _ in range(n_seq)]
    def init_with_sim_depths(self):
        for e in self.e: e.sim_init_depths = True
    def inertial_meas(self, ts, g, a):
        for e in self.e: e.InertialMeas(int(ts), np.asarray(g, float), np.asarray(a, float))
    def visual_meas_pointcloud(self, ts, ids, xpd, tracker_only=False):
        if isinstance(ids, np.ndarray): ids, xpd = [ids] * self.n, [xpd] * self.n
        for e, i, x in zip(self.e, ids, xpd): e.VisualMeasPointCloud(int(ts), np.asarray(i), np.asarray(x).reshape(-1, 3))
    def visual_meas(self, ts, imgs, tracker_only=False):
        for e, im in zip(self.e, imgs): (e.VisualMeasTrackerOnly if tracker_only else e.VisualMeas)(int(ts), im)
    def gsb(self, s=0): return self.e[s].gsb().copy()
    def P(self, s=0): return self.e[s].P.copy()
    def Pstate(self, s=0): return self.e[s].P[:9, :9].copy()
    def now(self, s=0): return int(self.e[s].curr_time)
    def motion(self, s=0):
        X = self.e[s].X; return X.Vsb.copy(), X.bg.copy(), X.ba.copy(), X.Rsg.copy()
    def counters(self, s=0):
        e = self.e[s]
        return dict(num_instate_features=len(e.instate_features), num_instate_groups=sum(g.instate() for g in e.groups.values()), gauge_group=e.gauge_group,
                    num_mh_rejected=e.num_mh_rejected, num_tracker_failed=0, num_tracker_new_detections=0, vision_counter=e.vision_counter, imu_counter=e.imu_counter,
                    MeasurementUpdateInitialized=int(e.meas_update_initialized), VisionInitialized=int(e.vision_initialized), num_tracked=len(e.tracks), error=0)
    def tracked_features(self, s=0, max_n=4096):
        t = self.e[s].tracks
        return np.array([f.id for f in t], np.int32), np.array([f.xp() for f in t]).reshape(-1, 2), np.array([f.tstatus for f in t], np.int32)
    def instate_features(self, s=0):
        fs = self.e[s].instate_features
        return dict(ids=np.array([f.id for f in fs], np.int32), sinds=np.array([f.sind for f in fs], np.int32), ref_groups=np.array([f.ref.id for f in fs], np.int32),
                    Xs=np.zeros((len(fs), 3)), x=np.array([f.x for f in fs]).reshape(-1, 3))
    def instate_feature_table(self, s=0, n_output=-1):
        e = self.e[s]; lay = e.lay
        if n_output < 0:
            fs = list(e.instate_features)
        else:
            fs = sorted((f for f in e.features.values() if f.instate()), key=lambda f: f.slot)
            fs = sorted(fs, key=lambda f: np.linalg.norm(e.P[lay.foff(f.sind):lay.foff(f.sind) + 3, lay.foff(f.sind):lay.foff(f.sind) + 3]))[:n_output]
        def cov(f):
            o = lay.foff(f.sind); c = e.P[o:o + 3, o:o + 3]
            return [c[0, 0], c[0, 1], c[0, 2], c[1, 1], c[1, 2], c[2, 2]]
        k = len(fs)
        return dict(ids=np.array([f.id for f in fs], np.int32), sinds=np.array([f.sind for f in fs], np.int32), ref_groups=np.array([f.ref.id for f in fs], np.int32),
                    Xs=np.array([e.feature_Xs(f)[0] for f in fs]).reshape(k, 3),
                    Xc=np.array([E.unproject_logz(f.x)[0] for f in fs]).reshape(k, 3), xc=np.array([f.x for f in fs]).reshape(k, 3),
                    pred=np.array([f.pred for f in fs]).reshape(k, 2), meas=np.array([f.xp() for f in fs]).reshape(k, 2), cov=np.array([cov(f) for f in fs]).reshape(k, 6))
    def instate_group_table(self, s=0):
        e = self.e[s]; lay = e.lay
        gs = [g for g in e.grps_std(lambda g: g.instate())]
        k = len(gs)
        return dict(ids=np.array([g.id for g in gs], np.int32), sinds=np.array([g.sind for g in gs], np.int32),
                    pose=np.array([np.r_[quat(g.Rsb), g.Tsb] for g in gs]).reshape(k, 7),
                    cov=np.array([e.P[lay.goff(g.sind):lay.goff(g.sind) + 6, lay.goff(g.sind):lay.goff(g.sind) + 6] for g in gs]).reshape(k, 6, 6))
    def calibration(self, s=0):
        e = self.e[s]; c = e.cam
        return dict(Ca=e.Ca.copy(), Cg=e.Cg.copy(), td=0.0, intrinsics=np.array([c.fx, c.fy, c.cx, c.cy, 0, 0, 0, 0, 0.0]), distortion_type=int(c.model))
    def just_dropped(self, s=0, max_n=4096): return np.zeros(0, np.int32)
    def tracker_counters(self, s=0):
        e = self.e[s]
        return dict(num_tracker_outlier_rejected=e.num_outliers_rejected, num_tracker_failed_to_track=e.num_failed_to_track, num_tracker_new_detections=0, num_oneptransac_rejected=0)
    def scale_init_velocity(self, scale, s=0): self.e[s].X.Vsb = self.e[s].X.Vsb / scale
    def close(self): pass

pyxivo.Batch = FakeBatch
