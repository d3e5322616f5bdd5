#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native XIVO inner loop.

Metric (BASELINE.json): VIO frames/s on synthetic 640x480 + 200 Hz IMU streams, whole job.
Workload (BASELINE.json configs[1]): full VIO, 640x480 pinhole, 150 tracked features, EKF state
dim 89 (G=4, F=14), `--seqs` independent sequences per GPU advancing in lock-step (the filter is
sequential per stream; batching sequences is the only parallel axis — SURVEY.md §8e).  One "step" =
one frame (+ its 8 IMU samples) for every sequence on every GPU.

  python bench.py --gpus N --steps K --warmup W              # this repo (CUDA)
  python bench.py --impl reference --gpus N --steps K --warmup W   # reference CPU arithmetic, all host cores

Prints ONE JSON line (rank 0).  `value` = frames/s with the frames already resident in HBM;
`e2e` = the same through the estimator-level C ABI with pinned HOST frames (H2D inside the timed
region, pose read back every step).  `roofline` describes the kernel with the largest share of
device time (CUDA-event durations recorded by the library on its launch streams) in a third pass over
the next frames of the same streams, in which the batches are stepped one after another so that an
event pair measures the kernel and not the queueing behind other batches' kernels.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS, COLS, G, F = 480, 640, 4, 14
IMU_PER_FRAME = 8
FRAME_NS = 40_000_000
PREROLL_FRAMES = 12  # gravity init (stationary) + first detections, never timed
CAL_ROUNDS, CAL_STEPS = 2, 4  # frame-ingest calibration: per round and mode one settling step + CAL_STEPS timed steps (untimed region)
INGEST_MODES = {"zero_copy": 0, "copy_engine": 1}


def load_cfg():
    from xivo_b200 import sim

    return sim.load_cfg(os.path.join(ROOT, "xivo_b200", "cfg", "vio_640x480.json"))


ALL_CPUS = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None


def cpu_reference(cores, frames, skip):
    """Runs oracle/cpu_baseline.py in a fresh interpreter (no CUDA context there: it forks one worker
    per core) and returns its JSON."""
    cfg_path = os.path.join(ROOT, "xivo_b200", "cfg", "vio_640x480.json")
    if hasattr(os, "sched_setaffinity"):
        os.sched_setaffinity(0, ALL_CPUS)  # the library pins its driver threads; the CPU arm gets every allowed CPU
    r = subprocess.run([sys.executable, "-m", "oracle.cpu_baseline", cfg_path, str(cores), str(frames), str(skip), str(G), str(F)], cwd=ROOT,
                       capture_output=True, text=True, timeout=900)
    if r.returncode != 0:
        raise RuntimeError("cpu baseline failed: " + r.stderr[-2000:])
    return json.loads(r.stdout.strip().splitlines()[-1])


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf=d.get("bf16_tflops_sustained", 1400.0), src="MEASURED_PEAKS.json (sustained)")
    return dict(hbm=6650.0, tf=1590.0, src="fallback B200_PROFILING.md")


class ClockSampler:
    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.idx = gpu_index

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i].strip().lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons, samples=len(sm))


def pick_ingest(ms):
    """ms: {mode: [ms per step of each calibration round]} -> the mode with the best round; the current default
    (zero_copy) keeps the job unless the alternative is at least 3 % faster (below that it is noise)."""
    best = {k: min(v) for k, v in ms.items() if v}
    if "zero_copy" not in best:
        return min(best, key=best.get)
    alt = min(best, key=best.get)
    return alt if best[alt] < 0.97 * best["zero_copy"] else "zero_copy"


def calibration_frames(choice):
    """Frames the ingest calibration consumes (run_ours renders that many more)."""
    return CAL_ROUNDS * len(INGEST_MODES) * (1 + CAL_STEPS) if choice == "auto" else 0


def calibrate_ingest(choice, set_mode, run_step, sync, f):
    """Pinned host frames reach the device either by a gather kernel reading host memory over PCIe or by the copy engine
    (xivo_set_frame_ingest); same results, the faster one depends on the host's PCIe path.  Picks per rank, outside every timed
    region, from end-to-end steps of this very workload: per round and mode one settling step + CAL_STEPS timed steps.
    set_mode(int), run_step(frame index), sync() are the caller's; returns (mode name, {mode: [ms per step, ...]} or None, next frame)."""
    if choice != "auto":
        set_mode(INGEST_MODES[choice])
        return choice, None, f
    ms = {m: [] for m in INGEST_MODES}
    for _ in range(CAL_ROUNDS):
        for name, mode in INGEST_MODES.items():
            set_mode(mode)
            run_step(f)
            f += 1
            sync()
            t0 = time.perf_counter()
            for _ in range(CAL_STEPS):
                run_step(f)
                f += 1
            sync()
            ms[name].append((time.perf_counter() - t0) * 1e3 / CAL_STEPS)
    best = pick_ingest(ms)
    set_mode(INGEST_MODES[best])
    return best, {k: [round(x, 3) for x in v] for k, v in ms.items()}, f


def make_streams(cfg, n_streams, n_frames):
    """n_streams distinct synthetic sequences (seeds 0..), each n_frames frames + IMU."""
    from xivo_b200 import sim

    out = []
    for s in range(n_streams):
        msgs, _ = sim.image_stream(cfg, duration=n_frames * 0.04 + 1e-9, seed=s, channels=1, fast=True)
        frames = [p for k, _, p in msgs if k == "img"]
        imu = [(ts, p) for k, ts, p in msgs if k == "imu"]
        out.append((frames[:n_frames], imu))
    return out


def cpu_budget():
    """CPUs this process may use: affinity mask capped by the cgroup CPU quota (the GPU boxes run with one)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            n = min(n, max(1, int(float(q) / float(p))))
    except Exception:
        try:
            q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            p = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, int(q / p)))
        except Exception:
            pass
    return n


def cpu_baseline_entry(r, cores, frames, wall_s):
    """`cpu_baseline` object from oracle/cpu_baseline.py's result.  kind "reference": a frame costs the OpenCV tracker calls (cv2 LK + FAST
    on the image stream) plus the reference's OWN estimator (oracle/_ref/libxivo_ref_*.so, its unmodified sources) on a point-cloud
    stream of the same state size and track count; kind "port" (library absent): only the third-party numerics are timed (upper bound)."""
    upd_us = None
    if r.get("stage_share", {}).get("update") is not None and r.get("mean_frame_ms"):
        upd_us = r["stage_share"]["update"] * r["mean_frame_ms"] * 1e3  # Eigen 3.3.9 Joseph update (UpdateJosephForm's expression sequence), one core
        upd_us = round(upd_us, 2) if math.isfinite(upd_us) else None
    if r.get("fps_reference"):
        return dict(value=r["fps_reference"], ekf_update_us_per_frame=upd_us, unit="frames/s", cores=cores, kind="reference",
                    sample=(f"{cores} concurrent processes x {frames} frames: per frame cv2 LK+FAST on a synthetic 640x480 sequence ({r['tracker_ms']:.2f} ms) + the reference's own "
                            f"estimator library (propagation, ProcessTracks, sub-filters, Jacobians, gating, Joseph update, management) on a point-cloud stream with "
                            f"{r['ref_tracks']:.0f} tracks, state dim 89 ({r['ref_estimator_ms']:.2f} ms); {wall_s:.0f}s wall"),
                    numerics_only_value=r["fps"], numerics_only_note="cv2 LK+FAST + Eigen 3.3.9 gate/update only (kind port): upper bound on the reference", stage_share=r["stage_share"])
    return dict(value=r["fps"], ekf_update_us_per_frame=upd_us, unit="frames/s", cores=cores, kind="port",
                sample=f"{cores} concurrent synthetic 640x480 sequences x {frames} frames; timed: cv2 LK+FAST and Eigen-3.3.9 gate+update on the restated pipeline's inputs ({r['mean_frame_ms']:.2f} ms/frame/core, eigen={r['eigen']}, {wall_s:.0f}s)",
                stage_share=r["stage_share"])


def build_roofline(prof, K, peaks, seqs_per_launch, pass_ms):
    """Pure post-processing of the library's profile report (xivo_profile_report): per-kernel CUDA-event time and the
    algorithmic work attributed at the launch sites -> the `roofline` object of the JSON line (dominant kernel by device
    time; `achieved` = work per launch / average launch duration) and the host phase breakdown.  Unit-tested on CPU."""
    kern = {k: v for k, v in prof.items() if not k.startswith("_") and not k.startswith("host:")}
    host_phases = {k[5:]: round(v["ms"] / K, 4) for k, v in prof.items() if k.startswith("host:")}
    upd_ms = kern.get("ekf_gain", {}).get("ms", 0) + kern.get("ekf_cov", {}).get("ms", 0)
    merged = {k: dict(v) for k, v in kern.items() if k not in ("ekf_gain", "ekf_cov", "ekf_update")}
    if upd_ms:
        merged["ekf_update"] = dict(calls=kern.get("ekf_gain", {}).get("calls", 0), ms=upd_ms, work=kern.get("ekf_update", {}).get("work", 0))
    if not merged:
        return dict(kernel=None, bound=None, achieved=None, peak=None, unit=None, frac=None, traffic=None, kernels={}), host_phases
    tot_ms = sum(v["ms"] for v in merged.values()) or 1.0
    dom = max(merged, key=lambda k: merged[k]["ms"])
    d = merged[dom]
    bound = "tensor" if dom == "ekf_update" else "hbm"
    per_launch_s = max(d["ms"], 1e-9) * 1e-3 / max(d["calls"], 1)
    work_per_launch = d.get("work", 0) / max(d["calls"], 1)
    if bound == "hbm":
        achieved, peak, unit = work_per_launch / per_launch_s / 1e9, peaks["hbm"], "GB/s"
    else:
        achieved, peak, unit = work_per_launch / per_launch_s / 1e12, peaks["tf"], "TFLOP/s"
    traffic, traffic_src = None, None
    try:  # ncu-measured DRAM bytes per launch of that kernel (profiles/), scaled to this run's sequences per launch
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if dom in tj:
            traffic = tj[dom] * (seqs_per_launch / tj["sequences_per_launch"])
            traffic_src = "ncu dram__bytes_{read,write}.sum at %d sequences/launch (profiles/r01_traffic.json), scaled to %d" % (tj["sequences_per_launch"], seqs_per_launch)
    except Exception:
        pass
    per_kernel = {}
    for k_, v_ in merged.items():  # the same arithmetic for every kernel with algorithmic work attributed (csrc/estimator.cu add_work sites)
        if not v_.get("work") or not v_["ms"]:
            continue
        tens = k_ == "ekf_update"
        a_ = v_["work"] / (v_["ms"] * 1e-3) / (1e12 if tens else 1e9)
        per_kernel[k_] = dict(bound="tensor" if tens else "hbm", achieved=round(a_, 4), unit="TFLOP/s" if tens else "GB/s",
                              frac=round(a_ / (peaks["tf"] if tens else peaks["hbm"]), 6))
    # BASELINE.json's second headline figure: EKF measurement-update time per frame (gain + covariance kernels, one launch = one update of
    # every filter of a batch)
    upd = merged.get("ekf_update")
    ekf_update = None
    if upd and upd["calls"]:
        ekf_update = dict(us_per_frame=round(upd["ms"] * 1e3 / (upd["calls"] * seqs_per_launch), 4), us_per_launch=round(upd["ms"] * 1e3 / upd["calls"], 3),
                          filters_per_launch=seqs_per_launch, launches=upd["calls"], kernels="ekf_gain_kernel + ekf_cov_kernel (or ekf_cov_tc_kernel)")
    roofline = dict(kernel=dom, bound=bound, achieved=achieved, peak=peak, unit=unit, frac=achieved / peak, traffic=traffic, traffic_source=traffic_src, peak_source=peaks["src"],
                    per_kernel=per_kernel, ekf_update=ekf_update,
                    share_of_device_time=d["ms"] / tot_ms, launches=d["calls"], avg_launch_us=per_launch_s * 1e6,
                    kernels={k: dict(ms=round(v["ms"], 4), calls=v["calls"], share=round(v["ms"] / tot_ms, 4)) for k, v in merged.items()},
                    device_busy_frac=tot_ms / pass_ms, profiled_pass_ms_per_step=pass_ms / K,
                    attribution="third pass, batches stepped one after another (event durations = kernels, not queueing behind other batches)")
    return roofline, host_phases


def run_ours(args):
    # host CPUs: the library's worker pool (workpool.h) is shared by the NB batches of this process; each batch
    # also has one driver thread (the Python thread inside xivo_batch_step), so workers + drivers = CPU budget
    budget = max(1, cpu_budget() // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))))
    # one driver per batch needs a CPU of its own: with a small budget (e.g. a node quota shared by 8 ranks) run fewer batches
    args.batches = max(1, min(args.batches, budget // 2))
    os.environ.setdefault("XIVO_THREADS", str(max(1, budget - args.batches + 1 - args.cpu_headroom)))
    os.environ.setdefault("XIVO_DRIVERS", str(args.batches))
    os.environ.setdefault("XIVO_PIN_DRIVERS", "1")  # the batch driver threads are ours: let the library pin them next to its workers
    import torch

    from xivo_b200 import capi, pyxivo

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = load_cfg()
    cfg["covariance_update"] = args.cov_update  # "fp64" (default, exact parity) or "tf32x3" (tcgen05 downdate, fp32 accuracy)
    B, K, W = args.seqs, args.steps, args.warmup
    n_cal = calibration_frames(args.ingest)
    n_frames = PREROLL_FRAMES + 3 * (W + K) + 4 + n_cal
    log("rendering", min(B, args.streams), "streams x", n_frames, "frames")
    streams = make_streams(cfg, min(B, args.streams), n_frames)
    log("streams ready")
    S = len(streams)
    # pinned host copies (e2e pass) and device copies (device-resident pass)
    host = torch.empty((S, n_frames, ROWS, COLS), dtype=torch.uint8).pin_memory()
    for s, (frames, _) in enumerate(streams):
        for f, img in enumerate(frames):
            host[s, f] = torch.from_numpy(img)
    dev = host.cuda()
    hnp = host.numpy()
    fbytes = ROWS * COLS
    host_ptr = [[hnp[s, f].ctypes.data for f in range(n_frames)] for s in range(S)]
    dev_ptr = [[dev.data_ptr() + (s * n_frames + f) * fbytes for f in range(n_frames)] for s in range(S)]
    from xivo_b200 import replicas

    seq_stream = replicas.assign_streams(rank, world, B, S)
    # IMU arrays per step: (B,3)
    imu_ts = np.array([[ts for ts, _ in streams[s][1]] for s in range(S)], dtype=np.uint64)
    imu_g = np.array([[p[0] for _, p in streams[s][1]] for s in range(S)])
    imu_a = np.array([[p[1] for _, p in streams[s][1]] for s in range(S)])

    # NB independent batches per GPU, each with its own stream and driven by its own host thread
    # (ctypes releases the GIL): the host phases of one batch overlap the kernels of the others.
    NB = max(1, min(args.batches, B))
    sizes = [B // NB + (1 if i < B % NB else 0) for i in range(NB)]
    L = capi.lib()
    import ctypes as C
    from concurrent.futures import ThreadPoolExecutor

    L.xivo_ctx_stream.restype = C.c_void_p
    ctxs, bts, exts, idxs = [], [], [], []
    o = 0
    for nb in sizes:
        c = capi.Context(local)
        ctxs.append(c)
        bts.append(pyxivo.Batch(cfg, n_seq=nb, max_groups=G, max_features=F, ctx=c))
        exts.append(torch.cuda.ExternalStream(L.xivo_ctx_stream(c._h)))
        idxs.append(np.array(seq_stream[o : o + nb]))
        o += nb
    pool = ThreadPoolExecutor(NB)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)

    def step_one(i, f, device_resident):
        idx, nb = idxs[i], sizes[i]
        ks = slice(f * IMU_PER_FRAME, (f + 1) * IMU_PER_FRAME)
        its = np.ascontiguousarray(imu_ts[idx, ks].T)                 # (8, nb)
        ig = np.ascontiguousarray(imu_g[idx, ks].transpose(1, 0, 2))  # (8, nb, 3)
        ia = np.ascontiguousarray(imu_a[idx, ks].transpose(1, 0, 2))
        ts = np.full(nb, f * FRAME_NS, dtype=np.uint64)
        ptrs = (C.c_void_p * nb)(*[(dev_ptr if device_resident else host_ptr)[s][f] for s in idx])
        rc = L.xivo_batch_step(bts[i]._h, IMU_PER_FRAME, vp(its), vp(ig), vp(ia), vp(ts), ptrs, ROWS, COLS, 1, int(device_resident))
        if rc != 0:
            raise RuntimeError(L.xivo_last_error().decode())
        return bts[i].gsb(0)  # host read of the step's result (pose); the err/P_mm D2H happened inside the call

    def step(f, device_resident, serial=False):
        if NB == 1 or serial:
            return [step_one(i, f, device_resident) for i in range(NB)]
        return list(pool.map(lambda i: step_one(i, f, device_resident), range(NB)))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
            torch.cuda.synchronize()

    f = 0
    for _ in range(PREROLL_FRAMES):
        step(f, True)
        f += 1
    log("preroll done", bts[0].counters(0))

    def timed(device_resident, profile):
        nonlocal f
        for _ in range(W):
            step(f, device_resident)
            f += 1
        L.xivo_profile_reset()
        L.xivo_profile_enable(int(profile))
        launches0 = capi.launch_count()
        clk = ClockSampler(local)
        barrier()
        clk.start()
        e0 = [torch.cuda.Event(enable_timing=True) for _ in range(NB)]
        e1 = [torch.cuda.Event(enable_timing=True) for _ in range(NB)]
        t0 = time.perf_counter()
        for i in range(NB):
            e0[i].record(exts[i])
        ntracked = 0
        for _ in range(K):
            # the attribution pass steps the batches one after another: with several batches in flight a CUDA-event pair
            # around a launch also measures the time the launch queued behind other batches' kernels, which made the
            # per-kernel shares disagree with the ncu launch list (profiles/r01g_bench.json vs r01c_launch_shares.txt)
            step(f, device_resident, serial=bool(profile))
            ntracked += bts[0].counters(0)["num_tracked"]
            f += 1
        for i in range(NB):
            e1[i].record(exts[i])
        barrier()
        wall = time.perf_counter() - t0
        clocks = clk.stop()
        ms = max(e0[0].elapsed_time(e1[i]) for i in range(NB))  # first start -> last end, on the launch streams
        L.xivo_profile_enable(0)
        buf = C.create_string_buffer(1 << 16)
        L.xivo_profile_report(buf, len(buf))
        prof = json.loads(buf.value.decode())
        ms = replicas.max_over_ranks(ms, device="cuda")
        return dict(ms=ms, wall_ms=wall * 1e3, prof=prof, launches=capi.launch_count() - launches0, clocks=clocks, ntracked=ntracked / K)

    # three passes over consecutive frames of the same streams: the two measured ones run with the in-library
    # profiler off (its event records and locks cost ~1 ms/step); the third only attributes time to kernels
    r_dev = timed(True, 0)
    log("device-resident pass", r_dev["ms"], "ms")
    L.xivo_set_frame_ingest.restype = C.c_int
    ingest, ingest_cal, f = calibrate_ingest(args.ingest, L.xivo_set_frame_ingest, lambda k: step(k, False), torch.cuda.synchronize, f)
    log("frame ingest:", ingest, ingest_cal)
    r_e2e = timed(False, 0)
    log("e2e pass", r_e2e["ms"], "ms")
    r_prof = timed(not args.profile_e2e, args.profile_level)
    log("profiled pass", r_prof["ms"], "ms")
    frames_total = world * B * K
    value = frames_total / (r_dev["ms"] * 1e-3)
    e2e = frames_total / (r_e2e["ms"] * 1e-3)

    peaks = measured_peaks()
    roofline, host_phases = build_roofline(r_prof["prof"], K, peaks, sizes[0], r_prof["ms"])

    out = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = min(cpu_budget(), args.cpu_cores) if args.cpu_cores else cpu_budget()  # CPUs the cgroup quota lets us run concurrently
            t0 = time.time()
            log("cpu baseline on", cores, "cores")
            r = cpu_reference(cores, 80, 14)
            cpu = cpu_baseline_entry(r, cores, 80, time.time() - t0)
        out = dict(metric="VIO frames/sec (640x480 synthetic + 200 Hz IMU)", value=value, unit="frames/s", n_gpus=world, steps=K, warmup=W,
                   ms_per_step=r_dev["ms"] / K, higher_is_better=True, scaling="weak", vs_baseline=None,
                   dtype="f64" if args.cov_update == "fp64" else "f64 state, 3xTF32 tensor-core covariance downdate", data="synthetic",
                   config=dict(covariance_update=args.cov_update, workload="BASELINE configs[1]: full VIO 640x480 pinhole + 200 Hz IMU, 150 tracked features, state dim 89 (G=4,F=14)",
                               sequences_per_gpu=B, batches_per_gpu=NB, host_threads=int(os.environ.get("XIVO_THREADS", "0")) or None, host_cpu_budget=budget, distinct_streams=S, frames_per_step=world * B, channels=1,
                               l2_policy="inputs larger than L2 are not needed: every step reads a new frame set (B x 307 KB) from the stream buffers; covariance/pyramids are the resident state by design",
                               message_buffer_size=cfg.get("message_buffer_size", 10), frame_ingest=ingest,
                               frame_ingest_calibration_ms_per_step=ingest_cal),
                   e2e=dict(value=e2e, unit="frames/s", h2d_bytes_per_step=r_e2e["prof"]["_h2d_bytes"] / K if r_e2e["prof"]["_h2d_bytes"] else world * B * fbytes,
                            d2h_bytes_per_step=(r_e2e["prof"]["_d2h_bytes"] / K) if r_e2e["prof"]["_d2h_bytes"] else None, ms_per_step=r_e2e["ms"] / K),
                   gpu_launches=r_dev["launches"], clocks=r_dev["clocks"], roofline=roofline, cpu_baseline=cpu,
                   tracked_features_mean=r_dev["ntracked"], wall_ms_per_step=r_dev["wall_ms"] / K, host_phase_ms_per_step=host_phases)
        print(json.dumps(out))
    pool.shutdown()
    for b_ in bts:
        b_.close()
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cfg = load_cfg()
    cores = min(cpu_budget(), args.cpu_cores) if args.cpu_cores else cpu_budget()  # CPUs the cgroup quota lets us run concurrently
    K, W = args.steps, args.warmup
    K_eff = min(K, 60)  # bounded sample: the restated pipeline around the timed numerics is Python
    t0 = time.time()
    r = cpu_reference(cores, K_eff, PREROLL_FRAMES + W)
    cb = cpu_baseline_entry(r, cores, K_eff, time.time() - t0)
    ms_per_step = 1e3 * cores / cb["value"]  # one step = one frame on each of `cores` concurrent sequences
    out = dict(impl="reference", metric="VIO frames/sec (640x480 synthetic + 200 Hz IMU)", value=cb["value"], unit="frames/s", n_gpus=args.gpus,
               steps=K_eff, warmup=W, ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
               config=dict(workload="BASELINE configs[1]: full VIO 640x480 pinhole + 200 Hz IMU, 150 tracked features, state dim 89 (G=4,F=14)",
                           sequences=cores, channels=1),
               cpu_baseline=cb,
               e2e=dict(value=cb["value"], unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-headroom", type=int, default=1, help="CPUs of the quota left to Python / CUDA helper threads")
    ap.add_argument("--profile-e2e", action="store_true", help="attribute kernel / host-phase time on the host-frame (e2e) path instead of the device-resident one")
    ap.add_argument("--profile-level", type=int, default=1, help="1: kernels + batch-level host phases, 2: + per-sequence host scopes")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seqs", type=int, default=512, help="independent sequences per GPU, split over --batches lock-step batches")
    ap.add_argument("--batches", type=int, default=8, help="independent lock-step batches per GPU, one driver thread each; their host phases share the library's worker pool")
    ap.add_argument("--streams", type=int, default=4, help="distinct synthetic input streams shared by the sequences")
    ap.add_argument("--cpu-cores", type=int, default=0)
    ap.add_argument("--cov-update", default="fp64", choices=["fp64", "tf32x3"], help="arithmetic of the covariance downdate (tf32x3 = tcgen05 tensor cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ingest", default="auto", choices=["auto"] + list(INGEST_MODES), help="how pinned host frames reach the device (e2e pass); auto = calibrate both before the timed region")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
