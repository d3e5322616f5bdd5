"""Runs the REFERENCE's own estimator (oracle/_ref/libxivo_ref_G<g>_F<f>.so, built by oracle/build_ref.py from the unmodified
sources under /root/reference) on a point-cloud-world message stream and returns / dumps its trajectory.  The reference keeps
process-wide singletons, so one estimator per process: use `run_subprocess` from tests.  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def lib_path(G: int, F: int, timing: bool = False) -> str:
    """timing=True: the -march=x86-64-v4 build when it exists and this CPU has AVX-512 (what the reference's -march=native gives
    on such a host); the parity pin always uses the x86-64-v3 build."""
    if timing:
        p4 = os.path.join(HERE, "_ref", f"libxivo_ref_G{G}_F{F}_v4.so")
        try:
            flags = open("/proc/cpuinfo").read()
        except OSError:
            flags = ""
        if os.path.exists(p4) and all(f in flags for f in ("avx512f", "avx512dq", "avx512bw", "avx512vl", "avx512cd")):
            return p4
    return os.path.join(HERE, "_ref", f"libxivo_ref_G{G}_F{F}.so")


def available(G: int, F: int) -> bool:
    return os.path.exists(lib_path(G, F))


def run(cfg: dict, msgs, G: int, F: int, sim_depths: bool = True, lib_file: str | None = None, on_visual=None):
    """msgs: [(kind, ts_ns, payload)] with 'imu' -> (gyro, accel), 'pc' -> (ids, xp_depth n x 3).  Returns a dict of arrays
    sampled after every visual message: gsb (k x 3 x 4), ts, n_instate, gauge, instate ids (list), and the final P."""
    L = C.CDLL(lib_file or lib_path(G, F))
    L.ref_error.restype = C.c_char_p
    L.ref_time_ns.restype = C.c_ulonglong
    assert L.ref_max_groups() == G and L.ref_max_features() == F
    if L.ref_create(json.dumps(cfg).encode()) != 0:
        raise RuntimeError("reference CreateSystem failed: " + L.ref_error().decode())
    if sim_depths:
        L.ref_init_with_sim_depths()
    N = L.ref_state_dim()
    out = dict(gsb=[], ts=[], n_instate=[], gauge=[], ids=[], vel=[])
    g12, cnt = np.zeros(12), (C.c_int * 6)()
    ids_buf, sind_buf = (C.c_int * 256)(), (C.c_int * 256)()
    v, bg, ba = np.zeros(3), np.zeros(3), np.zeros(3)
    for kind, ts, p in msgs:
        if kind == "imu":
            g, a = np.ascontiguousarray(p[0], dtype=np.float64), np.ascontiguousarray(p[1], dtype=np.float64)
            if L.ref_inertial(C.c_ulonglong(ts), g.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p)) != 0:
                raise RuntimeError(L.ref_error().decode())
        else:
            ids = np.ascontiguousarray(p[0], dtype=np.int32)
            xpd = np.ascontiguousarray(p[1], dtype=np.float64)
            if L.ref_visual_pointcloud(C.c_ulonglong(ts), len(ids), ids.ctypes.data_as(C.c_void_p), xpd.ctypes.data_as(C.c_void_p)) != 0:
                raise RuntimeError(L.ref_error().decode())
            L.ref_gsb(g12.ctypes.data_as(C.c_void_p))
            L.ref_counters(cnt)
            n = L.ref_instate_feature_ids(ids_buf, sind_buf, 256)
            L.ref_motion(v.ctypes.data_as(C.c_void_p), bg.ctypes.data_as(C.c_void_p), ba.ctypes.data_as(C.c_void_p))
            out["gsb"].append(g12.reshape(3, 4).copy())
            out["ts"].append(L.ref_time_ns())
            out["n_instate"].append(cnt[0])
            out["gauge"].append(cnt[2])
            out["ids"].append(sorted(ids_buf[i] for i in range(min(n, 256))))
            out["vel"].append(v.copy())
            if on_visual is not None:
                on_visual()
    P = np.zeros((N, N))
    L.ref_P(P.ctypes.data_as(C.c_void_p))
    out["P"] = P
    if hasattr(L, "ref_feature_table"):
        out["acc"] = accessor_dump(L)
    out["gsb"], out["ts"], out["vel"] = np.array(out["gsb"]), np.array(out["ts"], dtype=np.uint64), np.array(out["vel"])
    out["n_instate"], out["gauge"] = np.array(out["n_instate"]), np.array(out["gauge"])
    return out


def accessor_dump(L) -> dict:
    """The read-back surface of pybind11/pyxivo.cpp:332-398 as the reference returns it right now (boundary-parity fixtures):
    per-feature tables through the no-argument overloads (`all.*`) and the (int n_output) overloads for n = 5 and n = 50
    (`top5.*`, `top50.*`; only the rows the reference initialises), group tables, calibration, counters."""
    out = {}
    M = 256
    ids, sinds, refs = (C.c_int * M)(), (C.c_int * M)(), (C.c_int * M)()
    bufs = {k: np.zeros((M, w)) for k, w in (("Xs", 3), ("Xc", 3), ("xc", 3), ("pred", 2), ("meas", 2), ("cov", 6))}
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)

    def table(n_output):
        rows = L.ref_feature_table(n_output, ids, sinds, refs, ptr(bufs["Xs"]), ptr(bufs["Xc"]), ptr(bufs["xc"]), ptr(bufs["pred"]), ptr(bufs["meas"]), ptr(bufs["cov"]), M)
        return rows

    count = table(0)  # max(#in-state features in the graph, 0)
    for tag, n in (("all", -1), ("top5", 5), ("top50", 50)):
        rows = table(n)
        k = rows if n < 0 else min(n, count)
        out[f"{tag}.rows"] = np.array(rows)
        out[f"{tag}.ids"], out[f"{tag}.sinds"], out[f"{tag}.refs"] = np.array(ids[:k]), np.array(sinds[:k]), np.array(refs[:k])
        for name, b in bufs.items():
            out[f"{tag}.{name}"] = b[:k].copy()
    gid, gs, pose, cov = (C.c_int * 64)(), (C.c_int * 64)(), np.zeros((64, 7)), np.zeros((64, 21))
    ng = L.ref_group_table(gid, gs, ptr(pose), ptr(cov), 64)
    out["groups.ids"], out["groups.sinds"], out["groups.pose"] = np.array(gid[:ng]), np.array(gs[:ng]), pose[:ng].copy()
    out["groups.cov6"] = cov[:ng, :6].copy()  # the reference initialises only the first six columns (estimator_accessors.cpp, `cnt = 0` inside the row loop)
    Ca, Cg, Rsg, intr, Ps, td, dt = np.zeros(9), np.zeros(9), np.zeros(9), np.zeros(9), np.zeros(81), C.c_double(), C.c_int()
    L.ref_calibration(ptr(Ca), ptr(Cg), C.byref(td), ptr(Rsg), ptr(intr), C.byref(dt), ptr(Ps))
    out.update({"Ca": Ca.reshape(3, 3), "Cg": Cg.reshape(3, 3), "td": np.array(td.value), "Rsg": Rsg.reshape(3, 3), "intrinsics": intr,
                "distortion_type": np.array(dt.value), "Pstate": Ps.reshape(9, 9)})
    jd = (C.c_int * 512)()
    nj = L.ref_just_dropped(jd, 512)
    out["just_dropped"] = np.array(jd[:nj], dtype=np.int64)
    tc = (C.c_int * 4)()
    L.ref_tracker_counters(tc)
    out["tracker_counters"] = np.array(tc[:])
    tid, txy = (C.c_int * 1024)(), np.zeros((1024, 2))
    nt = L.ref_tracked_features(tid, ptr(txy), 1024)
    out["tracked.ids"], out["tracked.xy"] = np.array(tid[:nt]), txy[:nt].copy()
    return out


def _main():
    """python -m oracle.ref_runner <cfg.json> G F duration seed sim_depths out.npz — renders the PCW stream with xivo_b200.sim"""
    sys.path.insert(0, ROOT)
    from xivo_b200 import sim

    cfg_path, G, F, dur, seed, simd, outp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), sys.argv[7]
    cfg = sim.load_cfg(cfg_path)
    if len(sys.argv) > 8 and sys.argv[8] not in ("", "null"):
        cfg.update(json.loads(sys.argv[8]))
    msgs, traj = sim.pcw_stream(cfg, duration=dur, seed=seed)
    off = int(sys.argv[9]) if len(sys.argv) > 9 else 0  # vision stamps shifted by `off` ns: 0 keeps the IMU/vision timestamp ties
    msgs = [(k, ts + (off if k == "pc" else 0), p) for k, ts, p in msgs]
    msgs.sort(key=lambda m: (m[1], 0 if m[0] == "imu" else 1))
    r = run(cfg, msgs, G, F, bool(simd))
    ids = np.full((len(r["ids"]), max(1, max(len(x) for x in r["ids"]))), -1, dtype=np.int64)
    for i, x in enumerate(r["ids"]):
        ids[i, : len(x)] = x
    acc = {"acc." + k: v for k, v in r.get("acc", {}).items()}
    np.savez_compressed(outp, gsb=r["gsb"], ts=r["ts"], n_instate=r["n_instate"], gauge=r["gauge"], ids=ids, P=r["P"], vel=r["vel"],
                        truth=np.array([traj.pos(t * 1e-9) for t in r["ts"]]), **acc)


def run_subprocess(cfg_path, G, F, duration, seed, sim_depths, out_npz, overrides=None, pc_offset_ns=0):
    cmd = [sys.executable, "-m", "oracle.ref_runner", cfg_path, str(G), str(F), str(duration), str(seed), str(int(sim_depths)), out_npz,
           json.dumps(overrides) if overrides else "null", str(int(pc_offset_ns))]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference run failed:\n" + r.stdout[-2000:] + r.stderr[-4000:])
    return np.load(out_npz)


if __name__ == "__main__":
    _main()
