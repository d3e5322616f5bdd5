"""oracle/estimator_oracle.py — TEST INFRASTRUCTURE ONLY.

CPU (numpy fp64) restatement of the reference's per-frame pipeline: message heap, IMU propagation,
tracker bookkeeping, UpdateStep (ProcessTracks, SelectAndAddNewFeatures, Jacobians, MH gating,
FilterUpdate in Joseph form, AbsorbError, group/feature management).  It follows
/root/reference/src/{estimator,manager,update,tracker,graph,graphbase,mm,feature,options}.cpp with a
dense N x N covariance exactly like the reference (no device, no batching).  The tracker arithmetic
comes from oracle/tracker_oracle.c (pinned on cv2).

Parity status: PINNED on the reference's own estimator for the point-cloud path.  oracle/build_ref.py compiles the reference's
unmodified estimator sources into oracle/_ref/libxivo_ref_*.so (OpenCV / glog replaced by type-only header shims), and
tests/test_reference_pin.py requires this oracle to reproduce its trajectories: identical in-state id tables, gauge group and clock after
every frame, pose within 1e-11, covariance within 1e-12 relative (measured 3e-15 / 1e-18) on 7 sequences, incl. the orders that are
properties of libstdc++ (unordered_map iteration, unstable std::sort, heap ties — oracle/stdorder.py).  Golden copies:
tests/golden/reference_pcw.npz.  The image path (tracker half) is pinned through oracle/tracker_oracle.c <-> cv2 only.
Remaining convention shared with the product where the reference is not reproducible: a deterministic rotation instead of the
random_device-seeded shuffle of collinear gauge candidates (graph.cpp:337).
"""
from __future__ import annotations

import heapq
import copy
import math

import numpy as np

from . import ekf_oracle as E
from . import stdorder as SO
from . import tracker_oracle as T

CREATED, TRACKED, DROPPED = 0, 1, 2
F_CREATED, F_INITIALIZING, F_READY, F_INSTATE, F_REJECTED, F_REJ_TRACKER, F_NULLREFED, F_GAUGE = range(8)
G_CREATED, G_INSTATE, G_FLOATING, G_GAUGE = range(4)


class Group:
    def __init__(self, slot):
        self.slot = slot
        self.id = self.sind = -1
        self.lifetime = 0
        self.status = G_CREATED
        self.Rsb, self.Tsb = np.eye(3), np.zeros(3)

    def instate(self):
        return self.status in (G_INSTATE, G_GAUGE)


class Feature:
    def __init__(self, slot):
        self.slot = slot
        self.reset(-1, 0, 0)

    def reset(self, fid, x, y):
        self.id, self.sind, self.lifetime, self.init_counter = fid, -1, 0, 0
        self.status, self.tstatus, self.ref = F_CREATED, CREATED, None
        self.x = np.array([x, y, 2.0])
        self.P = np.zeros((3, 3))
        self.pred = np.array([-1.0, -1.0])
        self.outlier_counter = 0.0
        self.tri_ok = False  # triangulation_successful_ (feature.cpp:80)
        self.track = [np.array([x, y], dtype=np.float64)]
        self.response = 0.0
        self.descriptor = None          # Feature::descriptor() (feature.h:50): 32 bytes, set at detection, replaced when `differential`
        self.kp0 = (float(x), float(y))  # Feature::keypoint().pt: the keypoint the feature was created from (SetKeypoint, never updated)

    def instate(self):
        return self.status in (F_INSTATE, F_GAUGE)

    def xp(self):
        return self.track[-1]

    def z(self):
        with np.errstate(all="ignore"):
            return float(np.exp(np.float64(self.x[2])))  # IEEE like the C library (a diverged log-depth gives inf, not an exception)


class Pool:
    """CircBufWithHash, src/mm.cpp:35-121 (USE_MAPPER off)."""

    def __init__(self, n, cls):
        self.items = [cls(i) for i in range(n)]
        self.init = [False] * n
        self.act = [False] * n
        self.n_init, self.search = 0, 0

    def get(self):
        n = len(self.items)
        if self.n_init < n:
            while True:
                if not self.init[self.search]:
                    self.init[self.search] = self.act[self.search] = True
                    self.n_init += 1
                    r = self.items[self.search]
                    self.search = (self.search + 1) % n
                    return r
                self.search = (self.search + 1) % n
        start = self.search
        while True:
            if not self.act[self.search]:
                self.act[self.search] = True
                r = self.items[self.search]
                self.search = (self.search + 1) % n
                return r
            self.search = (self.search + 1) % n
            if self.search == start:
                raise RuntimeError("out of slots in the memory manager")

    def deactivate(self, t):
        self.act[t.slot] = False

    def destroy(self, t):
        if self.init[t.slot]:
            self.n_init -= 1
        self.act[t.slot] = self.init[t.slot] = False


class _InsertionOrderedIds(dict):
    """Set of ids that remembers the insertion order (a re-inserted id keeps its first position, like a hash set)."""

    def add(self, k):
        self.setdefault(k, True)


def rot_of(v):
    v = np.asarray(v, dtype=np.float64)
    if v.size == 3:
        return E.so3_exp(v)
    return E.quat_normalize_rot(v.reshape(3, 3))


class EstimatorOracle:
    def __init__(self, cfg: dict, G=15, F=30, tracker_only=False, std_order=True):
        # std_order: reproduce the reference's libstdc++-defined orders (oracle/stdorder.py); False = ascending ids + stable sorts,
        # the deterministic stand-in used before the reference itself could be run (kept to measure what the orders are worth)
        self.std_order = std_order
        self.um_features = SO.StdUnorderedIntMap() if std_order else None
        self.um_groups = SO.StdUnorderedIntMap() if std_order else None
        self.um_obs = {}
        self.std_heap = None
        self.lay = E.Layout(G, F)
        N = self.lay.N
        self.tracker_only = tracker_only
        c = cfg
        self.simulation = c.get("simulation", False)
        self.method = c.get("integration_method", "unspecified")
        sf = c.get("subfilter", {})
        self.sub_Rtri = sf.get("visual_meas_std", 3.5) ** 2
        self.sub_mh = sf.get("MH_thresh", 5.991)
        self.sub_ready = sf.get("ready_steps", 5)
        ad = c.get("adaptive_initial_depth", {})
        self.adapt_w, self.adapt_life = ad.get("median_weight", 0.99), ad.get("minimum_feature_lifetime", 5)
        self.remove_outlier_counter = c.get("remove_outlier_counter", 10)
        self.group_degrees_fixed = c.get("group_degrees_fixed", 4)
        self.max_group_lifetime = c.get("max_group_lifetime", 1)
        ic = c["imu_calib"]
        self.Ca = np.array(ic["Car"], float).reshape(3, 3) @ np.diag(ic["Cas"])
        self.Cg = np.array(ic["Cgr"], float).reshape(3, 3) @ np.diag(ic["Cgs"])
        self.g = np.array(c["gravity"], float)
        Xj = c["X"]
        bg, ba = np.array(Xj["bg"], float), np.array(Xj["ba"], float)
        if c.get("imu_tk_convention", False):
            bg, ba = -self.Cg @ bg, -self.Ca @ ba
        Wsg = np.array(list(Xj["Wsg"])[:2] + [0.0], float)
        self.X = E.MotionState(rot_of(Xj["Wsb"]), np.array(Xj["Tsb"], float), np.array(Xj["Vsb"], float), bg, ba, rot_of(Xj["Wbc"]),
                               np.array(Xj["Tbc"], float), E.so3_exp(Wsg))
        Pj = c["P"]
        d = np.ones(N)
        d[0:3], d[3:6], d[6:9], d[9:12], d[12:15], d[15:18] = Pj["Wsb"], Pj["Tsb"], Pj["Vsb"], Pj["bg"], Pj["ba"], Pj["Wbc"]
        d[18:21] = Pj["Tbc"]
        d[21:23] = Pj["Wsg"]
        self.P = np.diag(d * d)  # identity elsewhere (estimator.cpp:258-302)
        self.err = np.zeros(N)
        Qm = c["Qmodel"]
        q = np.zeros(23)
        q[0:3], q[15:18], q[21:23] = Qm["Wsb"], Qm["Wbc"], Qm["Wsg"]
        self.Qmodel = np.diag(q * q)
        Qi = c["Qimu"]
        qi = np.concatenate([np.broadcast_to(np.array(Qi[k], float), (3,)) for k in ("gyro", "accel", "gyro_bias", "accel_bias")])
        self.Qimu = np.diag(qi * qi)
        self.R = c["visual_meas_std"] ** 2
        cj = c["camera_cfg"]
        model = {"pinhole": 0, "equidistant": 3}[cj["model"]]
        self.cam = E.Camera(model, cj["rows"], cj["cols"], cj["fx"], cj["fy"], cj["cx"], cj["cy"], tuple(cj.get("k0123", (0, 0, 0, 0))))
        fl = self.cam.focal_length()
        self.init_z = c["initial_z"]
        self.init_std = (c["initial_std_x"] / fl, c["initial_std_y"] / fl, c["initial_std_z"])
        self.min_z, self.max_z = c["min_depth"], c["max_depth"]
        # depth triangulation before the sub-filter (estimator.cpp:157-164, :356-358; jsoncpp: a missing number reads as 0)
        self.tri_pre = c.get("triangulate_pre_subfilter", False)
        tr = c.get("triangulation", {})
        self.tri_method = tr.get("method", "l1_angular")
        self.tri_zmin, self.tri_zmax = tr.get("zmin", 0.05), tr.get("zmax", 5.0)
        self.tri_max_theta = tr.get("max_theta_thresh", 0.1) * math.pi / 180
        self.tri_beta = tr.get("beta_thesh", 0.25) * math.pi / 180
        self.init_std_badtri = (c.get("initial_std_x_badtri", 0.0), c.get("initial_std_y_badtri", 0.0), c.get("initial_std_z_badtri", 0.0))
        self.num_good_tri = self.num_bad_tri = 0
        # depth refinement of in-state candidates (estimator.cpp:143-155; manager.cpp:387-395, :431-440, :504-519)
        self.use_depth_opt = c.get("use_depth_opt", False)
        do = c.get("depth_opt", {})
        self.depth_opt = dict(two_view=do.get("two_view", False), use_hessian=do.get("use_hessian", False), max_iters=do.get("max_iters", 5),
                              eps=do.get("eps", 1e-4), max_res_norm=do.get("max_res_norm", 2.0))
        self.num_refined = self.num_refine_failed = 0
        self.use_MH = c.get("use_MH_gating", True)
        # filter-level 1-point RANSAC (update.cpp:213-393, call site manager.cpp:642-656); keys: estimator.cpp:131-134
        self.use_1pt = bool(c.get("use_1pt_RANSAC", False))
        self.ransac_thresh = float(c.get("1pt_RANSAC_thresh", 5))
        self.ransac_prob = float(c.get("1pt_RANSAC_prob", 0.95))
        self.ransac_chi2 = float(c.get("1pt_RANSAC_Chi2", 5.89))
        self.num_oneptransac_rejected = 0
        self.min_inliers = c.get("min_inliers", 5)
        self.MH_thresh, self.MH_mult = c.get("MH_thresh", 5.991), c.get("MH_adjust_factor", 1.1)
        self.owner_cov_factor = c.get("filter_owner_change_cov_factor", 1.5)
        self.strict_steps = c.get("strict_criteria_timesteps", 5)
        self.n_gauge = c.get("num_gauge_xy_features", 3)
        self.collinear_thresh = c.get("collinear_cross_prod_thresh", 1e-3)
        self.max_sub_outlier = c.get("max_subfilter_outlier", 0.01)
        self.h0 = c.get("PrinceDormand", {}).get("stepsize", 0.002) if self.method == "PrinceDormand" else c.get("RK4", {}).get("stepsize", 0.002)
        mem = c.get("memory", {})
        self.fpool = Pool(mem.get("max_features", 256), Feature)
        self.gpool = Pool(mem.get("max_groups", 128), Group)
        self.msg_buf_size = c.get("message_buffer_size", 10)
        if self.simulation:
            self.gravity_init_counter, self.gravity_initialized = 0, True
        else:
            self.gravity_init_counter, self.gravity_initialized = c.get("gravity_init_counter", 20), False
        tj = c["tracker_cfg"]
        self.t_mask_size, self.t_margin = tj.get("mask_size", 15), tj.get("margin", 16)
        self.t_min, self.t_max = tj.get("num_features_min", 120), tj.get("num_features_max", 150)
        self.t_max_disp = tj.get("max_pixel_displacement", 64)
        klt = tj.get("KLT", {})
        self.klt = dict(win=klt.get("win_size", 15), max_level=klt.get("max_level", 4), max_iter=klt.get("max_iter", 15), eps=klt.get("eps", 0.01))
        # tracker-level outlier rejection by homography (tracker.cpp:131-150)
        self.do_outlier_rejection = tj.get("do_outlier_rejection", False)
        oj = tj.get("outlier_rejection", {})
        self.outlier_rejection = ({"RANSAC": 8, "LMEDS": 4}[oj.get("method", "RANSAC")], oj.get("RANSAC_reproj_thresh", 3.0), oj.get("RANSAC_max_iters", 2000),
                                  oj.get("confidence", 0.995))
        self.num_outliers_rejected = self.num_failed_to_track = 0
        # descriptor path (tracker.cpp:176-217): BRIEF-32 per track and frame, descriptor check, rescue of dropped tracks, MATCH tracker
        self.desc_thresh = int(tj.get("descriptor_distance_thresh", -1))
        self.extract_descriptor = bool(tj.get("extract_descriptor", False)) or self.desc_thresh > -1
        self.differential = bool(tj.get("differential", True))
        self.tracker_type = tj.get("tracker_type", "LK")
        self.match_dropped = bool(tj.get("match_dropped_tracks", False)) and self.tracker_type == "LK"
        if self.tracker_type == "MATCH" and not self.extract_descriptor:
            raise ValueError("Using a matcher-tracker requires extracting descriptors")
        if self.match_dropped and not self.extract_descriptor:
            raise ValueError("must extract descriptors in order to match dropped tracks")
        if self.extract_descriptor and tj.get("descriptor", "BRIEF") != "BRIEF":
            raise ValueError("only the BRIEF descriptor is restated")
        self.fast_thr = tj.get("FAST", {}).get("threshold", 5)
        self.fast_nms = tj.get("FAST", {}).get("nonmaxSuppression", True)
        # bookkeeping
        self.gsel, self.fsel = [False] * G, [False] * F
        self.features, self.groups = {}, {}
        self.feature_adj, self.group_adj, self.gauge_features = {}, {}, {}
        self.tracks = []
        self.needs_new_gauge = []
        self.gauge_group = -1
        self.feature_counter, self.group_counter = 10000, 0
        self.ids_to_depths, self.sim_init_depths = {}, False
        self.instate_features = []
        self.tracker_initialized = False
        self.mask, self.prev_img = None, None
        self.num_mh_rejected = 0
        self.vision_initialized = False
        self.imu_counter = self.vision_counter = 0
        self.gravity_buf = []
        self.last_time = self.curr_time = 0
        self.last_accel = self.curr_accel = self.last_gyro = self.curr_gyro = np.zeros(3)
        self.slope_accel = self.slope_gyro = np.zeros(3)
        self.buf, self.buf_init, self.seq = [], False, 0
        self.meas_update_initialized = False
        self.stage_timer = None  # bench.py's CPU-baseline legs install a StageTimer here

    # ---------------------------------------------------------------- public API
    def InertialMeas(self, ts, gyro, accel):
        self._push((ts, 0, (np.array(gyro, float), np.array(accel, float))))

    def VisualMeasPointCloud(self, ts, ids, xp_depth):
        self._push((ts, 3, (np.array(ids), np.array(xp_depth, float).reshape(-1, 3))))

    def VisualMeas(self, ts, img):
        self._push((ts, 1, np.ascontiguousarray(img)))

    def VisualMeasTrackerOnly(self, ts, img):
        self._push((ts, 2, np.ascontiguousarray(img)))

    def gsb(self):
        return np.column_stack([self.X.Rsb, self.X.Tsb])

    # ---------------------------------------------------------------- message heap (estimator.cpp:923-941)
    def _push(self, m):
        # MaintainBuffer (estimator.cpp:923-941).  The reference's comparator looks at the timestamp only, so the execution order
        # among equal timestamps is a property of libstdc++'s heap algorithms: std_order reproduces it through oracle/stdumap.cpp;
        # without it ties are broken by arrival order (IMU pushed before vision on equal ts, as pyxivo_pcw.py:111 orders them).
        if self.std_order:
            if self.std_heap is None:
                self.std_heap = SO.StdMessageHeap(self.msg_buf_size)
            due = self.std_heap.push(m[0], m)
            if due is not None:
                self._execute(due)
            return
        item = (m[0], self.seq, m)
        self.seq += 1
        self.buf.append(item)
        if not self.buf_init:
            if len(self.buf) >= self.msg_buf_size:
                heapq.heapify(self.buf)
                self.buf_init = True
        else:
            heapq.heapify(self.buf)
        if self.buf_init and len(self.buf) > self.msg_buf_size:
            _, _, msg = heapq.heappop(self.buf)
            self._execute(msg)

    def _execute(self, m):
        ts, kind, p = m
        if kind == 0:
            self.inertial_internal(ts, p[0], p[1])
        else:
            self.visual_internal(ts, kind, p)

    # ---------------------------------------------------------------- time / imu
    def good_timestamp(self, now):
        return now // 1000000 >= self.curr_time // 1000000

    def inertial_internal(self, ts, gyro, accel):
        if not self.good_timestamp(ts):
            return
        self.imu_counter += 1
        if not self.gravity_initialized:
            self.gravity_buf.append(accel)
            if self.simulation or len(self.gravity_buf) >= self.gravity_init_counter:
                if not self.simulation:
                    mean = np.mean(self.gravity_buf, axis=0)
                    ac = self.Ca @ mean - self.X.ba
                    a, b = -self.g / np.linalg.norm(self.g), ac / np.linalg.norm(ac)
                    axis = np.cross(a, b)
                    s, cs = np.linalg.norm(axis), float(a @ b)
                    W = axis * (math.atan2(s, cs) / s) if s > 1e-12 else np.zeros(3)
                    W[2] = 0
                    self.X.Rsg = E.so3_exp(W)
                self.last_time = ts
                self.curr_accel = self.last_accel = accel
                self.curr_gyro = self.last_gyro = gyro
                self.gravity_initialized = True
                self.gravity_buf = []
        elif self.vision_initialized:
            self.last_time, self.curr_time = self.curr_time, ts
            self.curr_accel, self.curr_gyro = accel, gyro
            self.propagate(False)

    def propagate(self, visual):
        dt = (self.curr_time - self.last_time) * 1e-9
        if dt == 0:
            return
        if not visual:
            self.slope_accel = (self.curr_accel - self.last_accel) / dt
            self.slope_gyro = (self.curr_gyro - self.last_gyro) / dt
            accel0, gyro0 = self.last_accel, self.last_gyro
            self.last_accel, self.last_gyro = self.curr_accel, self.curr_gyro
        else:
            accel0, gyro0 = self.last_accel, self.last_gyro
            self.last_accel = accel0 + self.slope_accel * dt
            self.last_gyro = gyro0 + self.slope_gyro * dt
        Phi, Pmm = E.integrate(self.method, self.X, self.P[:23, :23].copy(), gyro0, accel0, self.slope_gyro, self.slope_accel, dt, self.Cg,
                               self.Ca, self.g, self.Qimu, self.h0)
        E.apply_propagation(self.P, Phi, Pmm, self.Qmodel)

    # ---------------------------------------------------------------- graph
    def g_add_feature(self, f):
        if self.std_order:
            self.um_features.insert(f.id)
        self.features[f.id] = f
        self.feature_adj[f.id] = {}
        if self.std_order and self.use_depth_opt:  # the inner unordered_map<int, Vec2> of feature_adj_ (graphbase.h:49-51): its iteration order
            self.um_obs[f.id] = SO.StdUnorderedIntMap()  # is read by GetObservationsOf (graphbase.cpp:146-152) -> RefineDepth's two_view pick

    def g_add_group(self, g):
        if self.std_order:
            self.um_groups.insert(g.id)
        self.groups[g.id] = g
        self.group_adj[g.id] = set()
        self.gauge_features[g.id] = set()

    def g_remove_feature(self, f):
        if self.std_order:
            self.um_features.erase(f.id)
        del self.features[f.id]
        for gid in self.feature_adj[f.id]:
            self.group_adj[gid].discard(f.id)
        del self.feature_adj[f.id]
        self.um_obs.pop(f.id, None)
        if f.ref is not None:
            self.gauge_features.setdefault(f.ref.id, set()).discard(f.id)

    def g_remove_group(self, g):
        if self.std_order:
            self.um_groups.erase(g.id)
        del self.groups[g.id]
        for fid in self.group_adj[g.id]:
            if g.id in self.feature_adj[fid] and fid in self.um_obs:
                self.um_obs[fid].erase(g.id)
            self.feature_adj[fid].pop(g.id, None)
        del self.group_adj[g.id]
        self.gauge_features.pop(g.id, None)

    def g_link(self, f, g):
        self.group_adj[g.id].add(f.id)
        if g.id not in self.feature_adj[f.id] and f.id in self.um_obs:
            self.um_obs[f.id].insert(g.id)
        self.feature_adj[f.id].setdefault(g.id, f.xp().copy())

    def feats(self, pred=lambda f: True):
        return [self.features[k] for k in sorted(self.features) if pred(self.features[k])]

    def grps(self, pred=lambda g: True):
        return [self.groups[k] for k in sorted(self.groups) if pred(self.groups[k])]

    def feats_std(self, pred=lambda f: True):
        """GraphBase::GetFeaturesIf in the reference's container order (graphbase.cpp:124-133)."""
        if not self.std_order:
            return self.feats(pred)
        return [self.features[k] for k in self.um_features.keys() if pred(self.features[k])]

    def grps_std(self, pred=lambda g: True):
        if not self.std_order:
            return self.grps(pred)
        return [self.groups[k] for k in self.um_groups.keys() if pred(self.groups[k])]

    def sort_candidates(self, feats):
        """std::sort(..., Criteria::CandidateComparison) (options.cpp:35-60) on `feats` in their current order."""
        if self.std_order:
            return SO.std_sort_candidates(feats)
        return sorted(feats, key=self.cand_key)

    # ---------------------------------------------------------------- features
    def create_feature(self, x, y):
        f = self.fpool.get()
        f.reset(self.feature_counter, x, y)
        self.feature_counter += 1
        return f

    def gbc(self):
        return self.X.Rbc, self.X.Tbc

    def feature_Xs(self, f):
        Xc, J = E.unproject_logz(f.x)
        Rsc = f.ref.Rsb @ self.X.Rbc
        Tsc = f.ref.Rsb @ self.X.Tbc + f.ref.Tsb
        return Rsc @ Xc + Tsc, Rsc @ J

    def change_owner(self, f, nref):
        Rsc = nref.Rsb @ self.X.Rbc
        Tsc = nref.Rsb @ self.X.Tbc + nref.Tsb
        Xs, dXs_dx = self.feature_Xs(f)
        Xcn = Rsc.T @ (Xs - Tsc)
        if Xcn[2] < 0:
            return False
        xn, dxn = E.project_logz(Xcn)
        J = dxn @ (Rsc.T @ dXs_dx)
        f.x = xn
        f.P = J @ f.P @ J.T
        f.ref = nref
        return True

    # ---------------------------------------------------------------- state slots
    def remove_group_from_state(self, g):
        E.remove_group_from_state(self.lay, self.P, self.err, g.sind)
        self.gsel[g.sind] = False
        g.sind, g.status = -1, G_FLOATING

    def remove_feature_from_state(self, f):
        E.remove_feature_from_state(self.lay, self.P, self.err, f.sind)
        self.fsel[f.sind] = False
        f.sind = -1

    def add_group_to_state(self, g):
        idx = self.gsel.index(False)
        self.gsel[idx] = True
        g.sind, g.status = idx, G_INSTATE
        E.add_group_to_state(self.lay, self.P, self.err, idx)

    def add_feature_to_state(self, f):
        idx = self.fsel.index(False)
        self.fsel[idx] = True
        f.status, f.sind = F_INSTATE, idx
        E.add_feature_to_state(self.lay, self.P, idx, f.P)

    # ---------------------------------------------------------------- visual
    def visual_internal(self, ts, kind, payload):
        if not self.good_timestamp(ts):
            return
        pc = kind in (3, 4)
        if pc != self.simulation:
            raise ValueError("VisualMeas / VisualMeasPointCloud called in the wrong mode")
        self.vision_counter += 1
        if not self.vision_initialized:
            if self.gravity_initialized:
                self.curr_time = ts
                self.vision_initialized = True
        else:
            self.last_time, self.curr_time = self.curr_time, ts
        tracker_only = kind in (2, 4)
        if not tracker_only:
            if not self.vision_initialized:
                return
            self.propagate(True)
            self.predict_features()
        if pc:
            ids, xpd = payload
            if kind == 3:
                for i, fid in enumerate(ids):
                    self.ids_to_depths.setdefault(int(fid), float(xpd[i, 2]))
            self.tracker_update_pointcloud(ids, xpd)
        elif self.tracker_type == "MATCH":
            self.tracker_update_match(payload)
        else:
            self.tracker_update_lk(payload)
        if tracker_only:
            for f in [f for f in self.tracks if f.tstatus == DROPPED]:
                self.tracks.remove(f)
                self.fpool.destroy(f)
            if self.gauge_group == -1:
                self.switch_ref_group()
            return
        self.update_step()
        if self.gauge_group == -1:
            self.switch_ref_group()

    def predict_features(self):
        gsb = (self.X.Rsb, self.X.Tsb)
        for f in self.tracks:
            if f.ref is None:
                continue
            f.pred = E.predict_pixel(self.cam, f.x, (f.ref.Rsb, f.ref.Tsb), gsb, self.gbc())

    def tracker_update_pointcloud(self, ids, xpd):
        meas = {int(i): xpd[k, :2].copy() for k, i in enumerate(ids)}
        marked = {int(i): False for i in ids}
        dropped = 0
        for f in self.tracks:
            m = meas.get(f.id)
            if m is not None and np.linalg.norm(m - f.xp()) < self.t_max_disp:
                f.track.append(m)
                f.tstatus = TRACKED
                marked[f.id] = True
            else:
                f.tstatus = DROPPED
                dropped += 1
        n_add = self.t_max - len(self.tracks) + dropped
        for fid in ids:
            if n_add <= 0:
                break
            fid = int(fid)
            if not marked[fid]:
                f = self.create_feature(*meas[fid])
                f.id = fid
                self.tracks.append(f)
            n_add -= 1

    def tracker_update_lk(self, img):
        rows, cols = img.shape[:2]
        if self.mask is None:
            self.mask = T.Mask(rows, cols, self.t_margin, self.t_mask_size)
        if not self.tracker_initialized:
            self.mask.m[:] = 0
            self.mask.reset()
            self.prev_img = img
            self.detect_lk(img, self.t_max)
            self.tracker_initialized = True
            return
        self.mask.reset()
        if not self.tracks:
            self.tracker_initialized = False
            return
        p0 = np.array([f.xp() for f in self.tracks], np.float32)
        p1 = p0.copy()
        for i, f in enumerate(self.tracks):
            if f.pred[0] != -1 and f.pred[1] != -1:
                p1[i] = f.pred.astype(np.float32)
                f.pred = np.array([-1.0, -1.0])
        r1, st, _ = T.lk_track(self.prev_img, img, p0, p1, **self.klt)
        if self.stage_timer is not None:
            self.stage_timer.lk(self.prev_img, img, p0, p1, self.klt)
        st = np.array(st, np.uint8).copy()
        if self.extract_descriptor:  # tracker.cpp:530-565: descriptors at the tracked positions; keypoints the extractor drops (border) are skipped
            desc, dvalid = T.brief(img, r1)
            for i, f in enumerate(self.tracks):
                if not dvalid[i]:
                    continue
                if self.desc_thresh != -1:
                    if T.hamming(f.descriptor, desc[i]) > self.desc_thresh:
                        st[i] = 0  # enforce to be dropped
                    elif self.differential:
                        f.descriptor = desc[i].copy()
                elif self.differential:
                    f.descriptor = desc[i].copy()
        valid, status = 0, np.zeros(len(self.tracks), np.uint8)
        for i, f in enumerate(self.tracks):
            ok = bool(st[i])
            if ok:
                q = r1[i].astype(np.float64)
                if self.mask.valid(q[0], q[1]) and np.linalg.norm(f.xp() - q) < self.t_max_disp:
                    f.tstatus = TRACKED
                    f.track.append(q)
                    self.mask.mask_out(q[0], q[1])
                    valid += 1
                else:
                    ok = False
            status[i] = ok
        self.num_failed_to_track = int((status == 0).sum())
        if self.do_outlier_rejection:  # tracker.cpp:594-599: homography outliers lose their status AFTER their track and the mask were updated
            from oracle import homography_oracle as HO

            done, rej, status = HO.tracker_outlier_rejection_413(p0, r1, status, *self.outlier_rejection)
            self._or_done = bool(done)
            if done:
                self.num_outliers_rejected = rej  # (stale when OutlierRejection returned early, like the reference's member)
            valid -= self.num_outliers_rejected
        dropped = [f for i, f in enumerate(self.tracks) if not status[i]]
        if valid < self.t_min:
            # (check_homography = outlier rejection ran: CheckHomography, tracker.cpp:818-828, never applies H -- it compares the dropped
            # feature's creation keypoint with the new one)
            self.detect_lk(img, self.t_max - valid, dropped, self.do_outlier_rejection and getattr(self, "_or_done", False))
        for f in dropped:  # tracker.cpp:617-619: every newly dropped track, rescued or not (DetectLK worked on a copy of the vector)
            f.tstatus = DROPPED
        self.prev_img = img

    def detect_lk(self, img, num_to_add, newly_dropped=(), check_homography=False):
        xy, sc, _ = T.fast_detect(img, self.fast_thr, self.fast_nms)
        if self.stage_timer is not None:
            self.stage_timer.fast(img, self.fast_thr, self.fast_nms)
        if not self.extract_descriptor:
            for i in T.select_keypoints(self.mask, xy, sc, num_to_add):
                f = self.create_feature(float(xy[i][0]), float(xy[i][1]))
                f.response = float(sc[i])
                self.tracks.append(f)
            return
        # with descriptors (tracker.cpp:231-327): mask filter and sort as above, the extractor then drops the keypoints of the border band
        keep = [i for i in range(len(xy)) if self.mask.m[int(xy[i][1] + 0.5), int(xy[i][0] + 0.5)]]
        keep.sort(key=lambda i: (-int(sc[i]), int(xy[i][1]), int(xy[i][0])))
        desc, dvalid = T.brief(img, np.asarray(xy, np.float32)[keep].reshape(-1, 2))
        keep = [i for i, v in zip(keep, dvalid) if v]
        desc = desc[dvalid]
        matched = {}  # position in `keep` -> index into newly_dropped
        newly_dropped = list(newly_dropped)
        if self.match_dropped and newly_dropped and keep:
            qd = np.stack([f.descriptor for f in newly_dropped])
            for qi, ti, dist in T.bf_match_crosscheck(qd, desc):
                f = newly_dropped[qi]
                k = keep[ti]
                ok_desc = (dist < self.desc_thresh) if self.desc_thresh > 0 else True  # CheckDescriptorDistance
                ok_disp = float(np.linalg.norm(np.array([float(xy[k][0]), float(xy[k][1])]) - f.xp())) < self.t_max_disp
                ok_h = True
                if check_homography:  # CheckHomography: |creation keypoint - new keypoint| < reprojection threshold (H is never applied)
                    ok_h = float(np.linalg.norm(np.array(f.kp0, np.float32).astype(np.float64) - np.array([float(xy[k][0]), float(xy[k][1])]))) < self.outlier_rejection[1]
                if ok_desc and ok_disp and ok_h:
                    matched[ti] = qi
        for pos, i in enumerate(keep):
            x, y = float(xy[i][0]), float(xy[i][1])
            if self.mask.valid(x, y):
                if self.match_dropped and pos in matched:
                    f1 = newly_dropped[matched[pos]]
                    if self.differential:
                        f1.descriptor = desc[pos].copy()
                    f1.track.append(np.array([x, y]))
                    f1.tstatus = TRACKED  # ("potentially rescued": UpdateLK marks it DROPPED again right after)
                    self.mask.mask_out(x, y)
                    num_to_add -= 1
                    continue  # (skips the budget / response test below, like the reference's `continue`)
                f = self.create_feature(x, y)
                f.response = float(sc[i])
                f.descriptor = desc[pos].copy()
                self.tracks.append(f)
                self.mask.mask_out(x, y)
                num_to_add -= 1
            if num_to_add <= 0 or sc[i] < 5:
                break

    def tracker_update_match(self, img):
        """Tracker::UpdateMatch (tracker.cpp:341-460): detect without a mask, describe, cross-checked nearest-neighbour matching of the
        existing features' descriptors against the new keypoints, checks, unmatched features dropped, unmatched keypoints become features."""
        img = np.ascontiguousarray(img)
        self.num_new_detections = 0
        xy, sc, _ = T.fast_detect(img, self.fast_thr, self.fast_nms)
        order = sorted(range(len(xy)), key=lambda i: (-int(sc[i]), int(xy[i][1]), int(xy[i][0])))
        desc, dvalid = T.brief(img, np.asarray(xy, np.float32)[order].reshape(-1, 2))
        order = [i for i, v in zip(order, dvalid) if v]
        desc = desc[dvalid]
        feats = list(self.tracks)
        new_matched, feat_matched = [False] * len(order), [False] * len(feats)
        if self.tracker_initialized and feats and order:
            matches = T.bf_match_crosscheck(np.stack([f.descriptor for f in feats]), desc)
            mstat = np.zeros(len(matches), np.uint8)
            for m, (qi, ti, dist) in enumerate(matches):
                k = order[ti]
                ok_desc = (dist < self.desc_thresh) if self.desc_thresh > 0 else True
                ok_disp = float(np.linalg.norm(np.array([float(xy[k][0]), float(xy[k][1])]) - feats[qi].xp())) < self.t_max_disp
                mstat[m] = ok_desc and ok_disp
            self.num_failed_to_track = len(feats) - len(matches) + int((mstat == 0).sum())
            if self.do_outlier_rejection and matches:
                from oracle import homography_oracle as HO

                p0 = np.array([feats[qi].kp0 for qi, _, _ in matches], np.float32)
                p1 = np.array([[float(xy[order[ti]][0]), float(xy[order[ti]][1])] for _, ti, _ in matches], np.float32)
                done, rej, mstat = HO.tracker_outlier_rejection_413(p0, p1, mstat, *self.outlier_rejection)
                if done:
                    self.num_outliers_rejected = rej
            for m, (qi, ti, dist) in enumerate(matches):
                if not mstat[m]:
                    continue
                new_matched[ti] = feat_matched[qi] = True
                f, k = feats[qi], order[ti]
                f.track.append(np.array([float(xy[k][0]), float(xy[k][1])]))
                if self.differential:
                    f.descriptor = desc[ti].copy()
                f.tstatus = TRACKED
        elif self.tracker_initialized:
            self.num_failed_to_track = len(feats)
        dropped = 0
        for f, m in zip(feats, feat_matched):
            if not m:
                f.tstatus = DROPPED
                dropped += 1
        to_create = self.t_max - len(feats) + dropped
        for pos, i in enumerate(order):
            if to_create <= 0:
                break
            if not new_matched[pos]:
                f = self.create_feature(float(xy[i][0]), float(xy[i][1]))
                f.response = float(sc[i])
                f.descriptor = desc[pos].copy()
                self.tracks.append(f)
                self.num_new_detections += 1
                to_create -= 1
        self.tracker_initialized = True

    # ---------------------------------------------------------------- UpdateStep (manager.cpp:18-167)
    def candidate(self, f, strict):
        good = (f.status == F_READY or (not strict and f.status == F_INITIALIZING)) and f.outlier_counter < self.max_sub_outlier
        return good and self.min_z < f.z() < self.max_z

    @staticmethod
    def cand_key(f):  # CandidateComparison: (status desc, score desc) with score = -P(2,2); stable
        return (-f.status, f.P[2, 2])

    def update_step(self):
        lay = self.lay
        self.ransac_trace = None
        affected, new_features, self.inliers, in_update = _InsertionOrderedIds(), [], [], []
        for f in self.features.values():
            f.lifetime += 1
        for g in self.groups.values():
            g.lifetime += 1
        gsb = (self.X.Rsb, self.X.Tsb)
        # ProcessTracks (manager.cpp:171-250)
        keep = []
        for f in self.tracks:
            if f.tstatus == CREATED:
                new_features.append(f)
            elif f.instate() and f.tstatus == DROPPED:
                self.g_remove_feature(f)
                if f.status == F_GAUGE:
                    self.needs_new_gauge.append(f.ref)
                self.remove_feature_from_state(f)
                affected.add(f.ref.id)
                self.fpool.deactivate(f)
            elif not f.instate() and f.tstatus == DROPPED:
                self.g_remove_feature(f)
                self.fpool.destroy(f)
            elif f.instate() and f.tstatus == TRACKED:
                keep.append(f)
            else:
                f.init_counter += 1
                if self.tri_pre and len(f.track) == 2:  # manager.cpp:229-231
                    self.triangulate(f)
                f.x, f.P, f.outlier_counter = E.subfilter_update(self.cam, f.x, f.P, f.xp(), gsb, self.gbc(), (f.ref.Rsb, f.ref.Tsb),
                                                                self.sub_Rtri, self.sub_mh, f.outlier_counter)
                f.status = F_READY if f.init_counter > self.sub_ready else F_INITIALIZING
                if f.outlier_counter > self.remove_outlier_counter:
                    self.g_remove_feature(f)
                    self.fpool.destroy(f)
                else:
                    keep.append(f)
        self.tracks = keep
        self.affected = affected
        inst = self.feats(lambda f: f.instate())
        if len(inst) < lay.F:
            self.select_and_add(inst)
        inst = sorted(set(inst), key=lambda f: f.slot)  # MakePtrVectorUnique
        # Jacobians + gating
        Js, inns = {}, {}
        for f in inst:
            J, r, _ = E.feature_jacobian(lay, self.cam, self.X.Rsb, self.X.Tsb, self.X.Rbc, self.X.Tbc, f.ref.Rsb, f.ref.Tsb, f.x, f.xp(),
                                         f.ref.sind, f.sind)
            Js[f.id], inns[f.id] = J, r
        if inst:
            if self.use_MH and len(inst) > self.min_inliers:
                dist = [E.mh_distance(Js[f.id], self.P, inns[f.id], self.R) for f in inst]
                if self.stage_timer is not None:
                    self.stage_timer.gate(np.array([Js[f.id] for f in inst]), self.P, np.array([inns[f.id] for f in inst]), self.R)
                self.num_mh_rejected = 0
                thresh, inl, to_destroy = self.MH_thresh, [], []
                while len(inl) < self.min_inliers:
                    for f in inst:
                        if f.status != F_GAUGE:
                            f.status = F_INSTATE
                    inl, to_destroy = [], []
                    for f, d in zip(inst, dist):
                        if d < thresh:
                            inl.append(f)
                        else:
                            self.num_mh_rejected += 1
                            to_destroy.append(f)
                    thresh *= self.MH_mult
                for f in to_destroy:
                    if f.status == F_GAUGE:
                        self.needs_new_gauge.append(f.ref)
                    f.status = F_REJECTED
                    affected.add(f.ref.id)
                for f in to_destroy:
                    self.g_remove_feature(f)
                for f in to_destroy:
                    self.remove_feature_from_state(f)
                    self.fpool.destroy(f)
                    self.tracks.remove(f)
                self.inliers = inl
            else:
                self.inliers = list(inst)
        if self.use_1pt and inst:  # manager.cpp:642-656: made observable first, then the global test
            before = set(self.features)
            self.discard_affected_groups()
            self.find_new_gauge_features()
            self.tracks = [f for f in self.tracks if not (f.id in before and f.id not in self.features)]
            self.inliers = self.one_point_ransac([f for f in self.inliers if f.instate()], Js, inns)
        before = set(self.features)
        self.discard_affected_groups()
        self.find_new_gauge_features()
        self.tracks = [f for f in self.tracks if not (f.id in before and f.id not in self.features)]
        in_update = [f for f in self.inliers if f.instate()]
        if in_update:
            inst_groups = self.grps(lambda g: g.instate())
            M = 2 * len(in_update)
            H, inn = np.zeros((M, lay.N)), np.zeros(M)
            for i, f in enumerate(in_update):
                E.fill_jacobian_block(lay, H, 2 * i, Js[f.id], f.ref.sind, f.sind)
                inn[2 * i : 2 * i + 2] = inns[f.id]
            if self.stage_timer is not None:
                self.stage_timer.update(H, self.P, inn, np.full(M, self.R))
            self.P, err, _, _ = E.update_joseph(H, self.P, inn, np.full(M, self.R))
            self.absorb(err, inst_groups, in_update)
        self.instate_features = in_update
        self.meas_update_initialized = True
        g = self.gpool.get()
        g.__init__(g.slot)
        g.id = self.group_counter
        self.group_counter += 1
        g.Rsb, g.Tsb = self.X.Rsb.copy(), self.X.Tsb.copy()
        self.g_add_group(g)
        self.tracks = []
        for f in new_features:
            f.ref = g
            z0 = self.ids_to_depths[f.id] if self.sim_init_depths else self.init_z
            f.x[:2] = self.cam.unproject(f.xp())
            f.x[2] = math.log(z0)
            f.P = np.diag(np.array(self.init_std) ** 2)
            if self.tri_pre and not f.tri_ok:  # manager.cpp:585-586: takes precedence over the simulated depths
                f.x[2] = math.log(self.init_z)
                f.P = np.diag(np.array(self.init_std_badtri) ** 2)
            f.status = F_INITIALIZING
            self.g_add_feature(f)
            self.g_link(f, g)
            self.tracks.append(f)
        for f in self.feats_std(lambda f: f.tstatus == TRACKED):  # AssociateTrackedFeaturesWithGroup: GetFeaturesIf order (manager.cpp:608-610)
            self.g_link(f, g)
            self.tracks.append(f)
        self.adapt_initial_depth()
        # EnforceMaxGroupLifetime (manager.cpp:282-303)
        for gg in self.grps():
            if gg.lifetime > self.max_group_lifetime and not any(self.features[fid].ref is gg for fid in self.group_adj[gg.id]):
                self.g_remove_group(gg)
                self.gpool.deactivate(gg)

    def one_point_ransac(self, mh_inliers, Js, inns):
        """Estimator::OnePointRANSAC (update.cpp:213-393).  What the code does (as opposed to what its comments describe):
        * every "hypothesis" evaluates the same test |xp - Predict(gsb, gbc)| < 1pt_RANSAC_thresh on every MH inlier -- the sampled index is
          never used -- so the low-innovation set is a plain threshold test and the RNG only decides how often it is repeated;
        * all low: the list is returned unchanged;
        * otherwise: back up X, P (and the active features' / groups' states); zero the P rows / columns of the high-innovation features and
          of the active groups without a low-innovation feature (and of a temporary reference group if the gauge group has none); a Joseph
          update with the FULL Jacobian rows J() of the low-innovation features (not FillJacobianBlock's layout); AbsorbError, which at
          this point of Estimator::Update moves only the motion state (instate_groups_ and in_current_ekf_update_ are still empty,
          manager.cpp:24, :90-103); every high-innovation feature is re-linearised at that state and kept if r' (J P J' + R)^-1 r <
          1pt_RANSAC_Chi2, else rejected (REJECTED_BY_FILTER, destroyed, its group affected); X, P and the saved states are restored --
          P_ = P0_ also brings the rows of the just freed slots back (they stay until the slot is reused; FillCovarianceBlock clears
          them, feature.cpp:753-760) -- and the Jacobians of the survivors are recomputed at the restored state.
        The reference returns the survivors in the iteration order of a std::unordered_set<FeaturePtr> (hash of heap addresses); here they
        keep the order of `mh_inliers`.  That order only permutes the rows of H in the update that follows."""
        if not mh_inliers:
            return mh_inliers
        lay = self.lay
        gsb = (self.X.Rsb, self.X.Tsb)
        low = []
        for f in mh_inliers:
            pred = E.predict_pixel(self.cam, f.x, (f.ref.Rsb, f.ref.Tsb), gsb, self.gbc())
            low.append(bool(np.linalg.norm(np.asarray(f.xp()) - pred) < self.ransac_thresh))
        if all(low):
            return mh_inliers
        # what the device phases of the product compute, for the CPU twin test (tests/test_host_twin.py)
        self.ransac_trace = dict(diag=np.diag(self.P).copy(), low=[f.id for f, l in zip(mh_inliers, low) if l], high=[f.id for f, l in zip(mh_inliers, low) if not l],
                                 err=None, mh={})
        active_groups = []
        for f in mh_inliers:
            if f.ref not in active_groups:
                active_groups.append(f.ref)
        groups_low = [g for g in active_groups if any(l and f.ref is g for f, l in zip(mh_inliers, low))]
        X0, P0 = copy.deepcopy(self.X), self.P.copy()
        fx0 = {f.id: f.x.copy() for f in mh_inliers}
        g0 = {g.id: (g.Rsb.copy(), g.Tsb.copy()) for g in active_groups}
        if any(low):
            gauge = self.groups.get(self.gauge_group)
            if gauge is None or gauge not in groups_low:  # temporary reference group: FindNewRefGroup = smallest trace of the 6x6 block
                cov = lambda g: float(np.trace(self.P[lay.goff(g.sind) : lay.goff(g.sind) + 6, lay.goff(g.sind) : lay.goff(g.sind) + 6]))
                tmp = min(sorted(groups_low, key=lambda g: g.id), key=cov)
                o = lay.goff(tmp.sind)
                self.P[o : o + 6, :] = 0
                self.P[:, o : o + 6] = 0
            for f, l in zip(mh_inliers, low):
                if not l:
                    o = lay.foff(f.sind)
                    self.P[o : o + 3, :] = 0
                    self.P[:, o : o + 3] = 0
            for g in active_groups:
                if g not in groups_low:
                    o = lay.goff(g.sind)
                    self.P[o : o + 6, :] = 0
                    self.P[:, o : o + 6] = 0
            lows = [f for f, l in zip(mh_inliers, low) if l]
            H, inn = np.zeros((2 * len(lows), lay.N)), np.zeros(2 * len(lows))
            for i, f in enumerate(lows):
                H[2 * i : 2 * i + 2] = Js[f.id]
                inn[2 * i : 2 * i + 2] = inns[f.id]
            self.P, err, _, _ = E.update_joseph(H, self.P, inn, np.full(2 * len(lows), self.R))
            self.ransac_trace["err"] = err.copy()
            self.absorb(err, [], [])
        out = [f for f, l in zip(mh_inliers, low) if l]
        rescued, to_destroy = [], []
        self.num_oneptransac_rejected = 0
        for f, l in zip(mh_inliers, low):
            if l:
                continue
            J, r, _ = E.feature_jacobian(lay, self.cam, self.X.Rsb, self.X.Tsb, self.X.Rbc, self.X.Tbc, f.ref.Rsb, f.ref.Tsb, f.x, f.xp(), f.ref.sind, f.sind)
            d_hi = E.mh_distance(J, self.P, r, self.R)
            self.ransac_trace["mh"][f.id] = d_hi
            if d_hi < self.ransac_chi2:
                rescued.append(f)
            else:
                if f.status == F_GAUGE:
                    self.needs_new_gauge.append(f.ref)
                f.status = F_REJECTED
                to_destroy.append(f)
                self.num_oneptransac_rejected += 1
                self.affected.add(f.ref.id)
        self.destroy_features(to_destroy)
        self.X, self.P = X0, P0
        for f in mh_inliers:
            f.x = fx0[f.id]
        for g in active_groups:
            g.Rsb, g.Tsb = g0[g.id]
        survivors = [f for f in mh_inliers if f in out or f in rescued]
        for f in survivors:  # Jacobians of the survivors at the restored state (the destroyed ones are recomputed too in the reference: dead objects)
            J, r, _ = E.feature_jacobian(lay, self.cam, self.X.Rsb, self.X.Tsb, self.X.Rbc, self.X.Tbc, f.ref.Rsb, f.ref.Tsb, f.x, f.xp(), f.ref.sind, f.sind)
            Js[f.id], inns[f.id] = J, r
        return survivors

    def triangulate(self, f):
        """Feature::Triangulate (feature.cpp:686-751) from the first and the newest observation of the track."""
        xc1, xc2 = self.cam.unproject(f.track[0]), self.cam.unproject(f.track[-1])
        Rbc, Tbc = self.gbc()
        Rrc, Trc = f.ref.Rsb @ Rbc, f.ref.Rsb @ Tbc + f.ref.Tsb
        Rsc, Tsc = self.X.Rsb @ Rbc, self.X.Rsb @ Tbc + self.X.Tsb
        x = E.triangulate(self.tri_method, Rrc.T @ Rsc, Rrc.T @ (Tsc - Trc), xc1, xc2, self.tri_zmin, self.tri_zmax, self.tri_max_theta,
                          self.tri_beta)
        if x is None:
            self.num_bad_tri += 1
        else:
            f.x = x
            f.tri_ok = True
            self.num_good_tri += 1

    def adapt_initial_depth(self):
        """AdaptInitialDepth (manager.cpp:255-278): the "median" is the middle element of the UNSORTED depth list, in graph order."""
        depth = [f.z() for f in self.feats_std(lambda f: f.instate() or (f.status == F_READY and f.lifetime > self.adapt_life))]
        if depth:
            med = depth[len(depth) >> 1]
            if not (med < self.min_z or med > self.max_z):
                self.init_z = (1 - self.adapt_w) * self.init_z + self.adapt_w * med

    def absorb(self, err, inst_groups, in_update):
        X = self.X
        X.Rsb = X.Rsb @ E.so3_exp(err[0:3])
        X.Tsb = X.Tsb + err[3:6]
        X.Vsb = X.Vsb + err[6:9]
        X.bg = X.bg + err[9:12]
        X.ba = X.ba + err[12:15]
        X.Rbc = X.Rbc @ E.so3_exp(err[15:18])
        X.Tbc = X.Tbc + err[18:21]
        X.Rsg = X.Rsg @ E.so3_exp(np.array([err[21], err[22], 0.0]))
        X.counter += 1
        if X.counter % 50 == 0:
            X.Rsb, X.Rbc = E.quat_normalize_rot(X.Rsb), E.quat_normalize_rot(X.Rbc)
            w = E.so3_log(X.Rsg)
            w[2] = 0
            X.Rsg = E.so3_exp(w)
        for g in inst_groups:
            o = self.lay.goff(g.sind)
            g.Rsb = g.Rsb @ E.so3_exp(err[o : o + 3])
            g.Tsb = g.Tsb + err[o + 3 : o + 6]
        for f in in_update:
            o = self.lay.foff(f.sind)
            f.x = f.x + err[o : o + 3]
        self.err[:] = 0

    def select_and_add(self, inst):  # manager.cpp:332-355
        free_g = self.gsel.count(False)
        free_f = self.lay.F - len(inst)
        if self.n_gauge == 0:
            self.zero_gauge_add(inst)
        elif free_f < self.n_gauge or free_g == 0:
            self.add_within_groups(inst)
        else:
            self.add_group_of_features(inst, free_g)
            self.add_within_groups(inst)

    def add_within_groups(self, inst):
        strict = not (self.vision_counter < self.strict_steps)
        cands = sorted(self.feats(lambda f: self.candidate(f, strict) and f.ref.instate()), key=lambda f: f.slot)  # MakePtrVectorUnique
        cands = self.sort_candidates(cands)
        bad = []
        for f in cands:
            if len(inst) >= self.lay.F:
                break
            if self.use_depth_opt and len(self.feature_adj[f.id]) > 1 and not self.refine_depth(f):
                bad.append(f)
                continue
            inst.append(f)
            self.add_feature_to_state(f)
        self.destroy_features(bad)

    def observations_of(self, f):
        """Graph::GetObservationsOf (graphbase.cpp:146-152): (group, pixel) in the inner container's iteration order."""
        keys = self.um_obs[f.id].keys() if f.id in self.um_obs else sorted(self.feature_adj[f.id])
        return [(self.groups[g], self.feature_adj[f.id][g]) for g in keys]

    def refine_depth(self, f):
        obs = self.observations_of(f)
        views = [(g.Rsb, g.Tsb, xp) for g, xp in obs]
        ref_index = next((i for i, (g, _) in enumerate(obs) if g.id == f.ref.id), -1)
        o = self.depth_opt
        ok, f.x, f.P, f.Xs = E.refine_depth(self.cam, f.x, f.P, (f.ref.Rsb, f.ref.Tsb), self.gbc(), views, ref_index, o["two_view"], o["use_hessian"],
                                            o["max_iters"], o["eps"], o["max_res_norm"], self.sub_Rtri)
        self.num_refined += ok
        self.num_refine_failed += not ok
        return ok

    def destroy_features(self, feats):
        """Estimator::DestroyFeatures (estimator.cpp:1352-1360)."""
        for f in feats:
            self.g_remove_feature(f)
        for f in feats:
            if f.instate() or f.status == F_REJECTED:
                self.remove_feature_from_state(f)
            self.fpool.destroy(f)
            if f in self.tracks:
                self.tracks.remove(f)

    def zero_gauge_add(self, inst):
        free = self.gsel.count(False)
        strict = not (self.vision_counter < self.strict_steps)
        cands = sorted(self.feats(lambda f: self.candidate(f, strict)), key=lambda f: f.slot)
        cands = self.sort_candidates(cands)
        bad = []
        for f in cands:
            if len(inst) >= self.lay.F:
                break
            if self.use_depth_opt and len(self.feature_adj[f.id]) > 1 and not self.refine_depth(f):
                bad.append(f)
                continue
            if not f.ref.instate() and free <= 0:
                continue
            inst.append(f)
            self.add_feature_to_state(f)
            if not f.ref.instate():
                self.add_group_to_state(f.ref)
                self.needs_new_gauge.append(f.ref)
                free -= 1
        self.destroy_features(bad)

    def add_group_of_features(self, inst, free_g):
        to_add = self.lay.F - len(inst)
        n_owned = lambda g: sum(1 for f in self.features.values() if f.ref is g and f.status == F_READY)
        cands = self.grps_std(lambda g: g.status == G_CREATED and n_owned(g) >= self.n_gauge)  # GetInstateGroupCandidates
        cands = SO.std_sort_desc(cands, [n_owned(g) for g in cands]) if self.std_order else sorted(cands, key=lambda g: -n_owned(g))
        for g in cands:
            feats = self.sort_candidates(self.feats_std(lambda f: f.ref is g and f.status == F_READY))  # GetFeatureCandidatesOwnedBy + std::sort
            if self.use_depth_opt:
                # manager.cpp:504-536: only features that refine well are added; if fewer than the gauge count do, the failed ones are
                # destroyed and the group marked affected — and the group is added to the state all the same (fall-through at :553)
                good, bad = [], []
                for f in feats:
                    if len(self.feature_adj[f.id]) > 1:
                        (good if self.refine_depth(f) else bad).append(f)
                if len(good) >= self.n_gauge:
                    feats = good
                else:
                    self.destroy_features(bad)
                    self.affected.add(g.id)
                    feats = []
            for f in feats:
                self.add_feature_to_state(f)
                inst.append(f)
                to_add -= 1
                if to_add == 0:
                    break
            self.add_group_to_state(g)
            self.needs_new_gauge.append(g)
            free_g -= 1
            if to_add < self.n_gauge or free_g == 0:
                break

    def discard_affected_groups(self):  # manager.cpp:307-328
        # affected_groups_ is a std::unordered_set<GroupPtr> (estimator.h:364): with its few elements in distinct buckets libstdc++ links every
        # new node at the head of the list, so the walk meets the groups in REVERSE insertion order (exact for two groups, the common case
        # of more than one; with three or more and a bucket collision the reference's order depends on heap addresses).  The order matters
        # when a discarded group's features find their new owner in another affected group.
        for gid in reversed(list(self.affected)):
            g = self.groups.get(gid)
            if g is None:
                continue
            n_in = sum(1 for f in self.features.values() if f.ref is g and f.instate())
            if n_in < self.n_gauge or (self.n_gauge == 0 and n_in == 0):
                failed = []
                for fid in sorted(self.group_adj[gid]):
                    f = self.features[fid]
                    if f.ref is not g:
                        continue
                    nref = None
                    for ogid in sorted(self.feature_adj[fid]):
                        og = self.groups[ogid]
                        if ogid != gid and og.status == G_INSTATE:
                            nref = og
                            break
                    if nref is not None:
                        ok = self.change_owner(f, nref)
                        f.P = f.P * self.owner_cov_factor
                        if not ok:
                            failed.append(f)
                    else:
                        failed.append(f)
                for f in failed:
                    self.g_remove_feature(f)
                    if f.instate():
                        self.remove_feature_from_state(f)
                    self.fpool.deactivate(f)
                for f in failed:
                    f.status = F_NULLREFED
                if g.id == self.gauge_group:
                    self.gauge_group = -1
                self.g_remove_group(g)
                if g.instate():
                    self.remove_group_from_state(g)
                self.gpool.deactivate(g)
        self.affected = _InsertionOrderedIds()

    def find_new_gauge_features(self):  # update.cpp:35-47 + graph.cpp:276-361
        for g in self.needs_new_gauge:
            if self.groups.get(g.id) is not g:
                continue
            gf = self.gauge_features[g.id]
            num_to_find = self.n_gauge - len(gf)
            cands = self.feats_std(lambda f: f.status == F_INSTATE and f.ref is g)  # GetGaugeFeatureCandidates (graph.cpp:107-112)
            backup_c = list(cands)

            def fill(C):
                out = []
                for f in C[: max(0, min(num_to_find, len(C)))]:
                    gf.add(f.id)
                    out.append(f)
                return out

            found = []
            if not cands or num_to_find == 0:
                pass
            elif len(cands) <= num_to_find:
                found = fill(cands)
            else:
                backup = set(gf)
                for NT in range(10):
                    gf.clear()
                    gf.update(backup)
                    found = fill(cands)
                    if len(gf) >= 3:
                        pts = [E.unproject_logz(self.features[fid].x)[0] for fid in sorted(gf)]
                        v1 = pts[1] - pts[0]
                        col = all(np.linalg.norm(np.cross(v1, p - pts[0])) <= self.collinear_thresh for p in pts[2:])
                        if col:
                            cands = cands[1:] + cands[:1]
                            if NT == 9:
                                gf.clear()
                                gf.update(backup)
                                fill(backup_c)
                        else:
                            break
            for f in found:
                f.status = F_GAUGE
                E.fix_feature_xy(self.lay, self.P, f.sind)
        self.needs_new_gauge = []

    def switch_ref_group(self):  # estimator.cpp:1362-1407
        cands = self.grps(lambda g: g.instate())
        if not cands:
            return
        cov = lambda g: float(np.trace(self.P[self.lay.goff(g.sind) : self.lay.goff(g.sind) + 6, self.lay.goff(g.sind) : self.lay.goff(g.sind) + 6]))
        best = cands[0]
        for g in cands:
            if cov(g) < cov(best):
                best = g
        self.gauge_group = best.id
        best.status = G_GAUGE
        E.switch_ref_group_cov(self.lay, self.P, best.sind, self.group_degrees_fixed)
