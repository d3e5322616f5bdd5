"""GPU debug aid: product vs oracle on a 1-point RANSAC pin case, frame by frame (first pose / covariance divergence)."""
import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.filterwarnings("ignore")
import test_reference_pin as RP
from oracle.estimator_oracle import EstimatorOracle
from xivo_b200 import pyxivo, sim

case = sys.argv[1] if len(sys.argv) > 1 else "ransac_outliers_203"
name, G, F, duration, seed, sim_depths, over, offset = next(c for c in RP.CASES if c[0] == case)
cfg = sim.load_cfg(RP.CFG); cfg.update(over or {})
msgs, _ = RP.stream(cfg, duration, seed, offset)
est = EstimatorOracle(cfg, G=G, F=F); est.sim_init_depths = sim_depths
b = pyxivo.Batch(cfg, n_seq=1, max_groups=G, max_features=F)
if sim_depths: b.init_with_sim_depths()
k = 0
reported = 0
for kind, ts, p in msgs:
    if kind == "imu":
        est.InertialMeas(ts, p[0], p[1]); b.inertial_meas(ts, p[0], p[1])
    else:
        est.VisualMeasPointCloud(ts, p[0], p[1]); b.visual_meas_pointcloud(ts, p[0], p[1])
        dg = np.abs(b.gsb(0) - est.gsb()).max()
        tr = getattr(est, "ransac_trace", None)
        P = b.P(0)
        dP = np.abs(P - est.P)
        rel = dP.max() / np.abs(est.P).max()
        if (tr is not None and est.num_oneptransac_rejected > 0) or dg > 1e-9 or rel > 1e-9:
            rows = np.nonzero(dP.max(1) > 1e-9 * np.abs(est.P).max())[0]
            print(f"frame {k}: ransac={'yes low %s high %s rej %d' % (tr['low'], tr['high'], est.num_oneptransac_rejected) if tr else 'no'} |dgsb|={dg:.2e} |dP|rel={rel:.2e} rows differing {rows[:12].tolist()}{'...' if len(rows) > 12 else ''} ids_equal={sorted(b.instate_features(0)['ids'].tolist()) == sorted(f.id for f in est.instate_features)}")
            if dg > 1e-9 or rel > 1e-9:
                print("   all rows differing:", rows.tolist())
                print("   oracle in-state groups (id, sind, status):", sorted((g.id, g.sind, g.status) for g in est.groups.values() if g.instate()), "gauge", est.gauge_group)
                print("   oracle in-state features (id, sind, ref id, ref sind, status):", [(f.id, f.sind, f.ref.id, f.ref.sind, f.status) for f in sorted(est.instate_features, key=lambda f: f.sind)])
                pf = b.instate_features(0); pg = b.instate_groups(0)
                print("   product features ids/sinds/refs:", list(zip(pf["ids"].tolist(), pf["sinds"].tolist(), pf["refs"].tolist())))
                print("   product groups:", list(zip(pg["ids"].tolist(), pg["sinds"].tolist())), "gauge", b.counters(0)["gauge_group"])
                blk = lambda r0, r1: float(dP[r0:r1].max())
                print("   |dP| motion %.2e groups %.2e features %.2e ; |P| max %.2e" % (blk(0, 23), blk(23, 23 + 6 * G), blk(23 + 6 * G, dP.shape[0]), np.abs(est.P).max()))
                break
            reported += 1
            if reported > 10: break
        k += 1
