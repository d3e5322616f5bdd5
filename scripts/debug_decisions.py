"""GPU debug aid: the image pipeline with device-side vs host-side tracker decisions, side by side; prints the first frame where the
track lists differ and what is known about the feature in question."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xivo_b200 import pyxivo, sim

CFG = os.path.join(ROOT, "xivo_b200", "cfg")
cfg = sim.load_cfg(os.path.join(CFG, "vio_640x480.json"))
cfg["camera_cfg"].update(rows=240, cols=320, fx=137.5, fy=137.5, cx=160, cy=120)
cfg["tracker_cfg"].update(num_features_min=60, num_features_max=80)
msgs, traj = sim.image_stream(cfg, duration=1.6, channels=1, seed=1)


def make(host):
    os.environ["XIVO_HOST_TRACKER_DECISIONS"] = "1" if host else "0"
    b = pyxivo.Batch(cfg, n_seq=1, max_groups=4, max_features=14)
    # the flag is read when the first image arrives
    return b


bd, bh = make(False), None
os.environ["XIVO_HOST_TRACKER_DECISIONS"] = "0"
first = True
prev = None
nf = 0
for kind, ts, p in msgs:
    if kind == "imu":
        bd.inertial_meas(ts, p[0], p[1])
        if bh: bh.inertial_meas(ts, p[0], p[1])
        else:
            if first:
                os.environ["XIVO_HOST_TRACKER_DECISIONS"] = "1"
                bh = pyxivo.Batch(cfg, n_seq=1, max_groups=4, max_features=14)
                first = False
            bh.inertial_meas(ts, p[0], p[1])
    else:
        os.environ["XIVO_HOST_TRACKER_DECISIONS"] = "0"
        bd.visual_meas(ts, [p])
        os.environ["XIVO_HOST_TRACKER_DECISIONS"] = "1"
        bh.visual_meas(ts, [p])
        nf += 1
        idd, xyd, std = bd.tracked_features(0)
        idh, xyh, sth = bh.tracked_features(0)
        if idd.tolist() != idh.tolist() or not np.array_equal(xyd, xyh):
            print("frame", nf, "device", len(idd), "host", len(idh))
            sd, sh = set(idd.tolist()), set(idh.tolist())
            print("only device:", sorted(sd - sh), "only host:", sorted(sh - sd))
            for fid in sorted(sh - sd):
                i = idh.tolist().index(fid)
                print("  host keeps", fid, "at", xyh[i], "status", sth[i])
                if prev is not None and fid in prev[0].tolist():
                    j = prev[0].tolist().index(fid)
                    print("  previous position", prev[1][j], "index in list", j, "of", len(prev[0]))
                d = np.abs(xyh - xyh[i]).max(1)
                near = [(int(idh[k]), xyh[k].tolist(), k) for k in np.argsort(d)[:5]]
                print("  nearest tracks (host run):", near)
            common = [f for f in idd.tolist() if f in sh]
            dx = [np.abs(xyd[idd.tolist().index(f)] - xyh[idh.tolist().index(f)]).max() for f in common]
            print("max position difference over common tracks", max(dx) if dx else None)
            print("counters device", bd.counters(0), "\ncounters host", bh.counters(0))
            break
        prev = (idh.copy(), xyh.copy())
else:
    print("no difference in", nf, "frames")
