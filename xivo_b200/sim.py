"""Synthetic visual-inertial data: analytic trajectory -> IMU samples, a random
point-cloud world for the VisualMeasPointCloud path (the reference's own simulation entry,
scripts/pyxivo_pcw.py) and a textured-plane renderer for the image path.  numpy/scipy only.
Test/bench infrastructure — not part of the hot path."""
from __future__ import annotations

import dataclasses
import json
import re

import numpy as np
from scipy import ndimage

G_S = np.array([0.0, 0.0, -9.8])


def strip_json_comments(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = []
    for line in text.splitlines():
        i, in_str = 0, False
        while i < len(line):
            ch = line[i]
            if ch == '"' and (i == 0 or line[i - 1] != "\\"):
                in_str = not in_str
            if not in_str and line[i : i + 2] == "//":
                line = line[:i]
                break
            i += 1
        out.append(line)
    text = "\n".join(out)
    return re.sub(r",(\s*[}\]])", r"\1", text)


def load_cfg(text_or_path: str) -> dict:
    if "{" not in text_or_path:
        text_or_path = open(text_or_path).read()
    return json.loads(strip_json_comments(text_or_path))


def rodrigues(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th**2 * (K @ K)


@dataclasses.dataclass
class Trajectory:
    """Lissajous position + small oscillating attitude.  Starts at rest at the origin with identity
    attitude (p, v, theta all zero at t = 0 thanks to the (1 - cos) / sin^2 shaping)."""

    amp: np.ndarray = dataclasses.field(default_factory=lambda: np.array([0.6, 0.4, 0.2]))
    freq: np.ndarray = dataclasses.field(default_factory=lambda: np.array([0.5, 0.35, 0.25]))
    rot_amp: np.ndarray = dataclasses.field(default_factory=lambda: np.array([0.10, 0.08, 0.25]))
    rot_freq: np.ndarray = dataclasses.field(default_factory=lambda: np.array([0.3, 0.2, 0.15]))

    def pos(self, t):
        w = 2 * np.pi * self.freq
        return self.amp * (1 - np.cos(w * t))

    def vel(self, t):
        w = 2 * np.pi * self.freq
        return self.amp * w * np.sin(w * t)

    def acc(self, t):
        w = 2 * np.pi * self.freq
        return self.amp * w * w * np.cos(w * t)

    def theta(self, t):
        w = 2 * np.pi * self.rot_freq
        return self.rot_amp * np.sin(w * t) ** 2

    def R(self, t):
        return rodrigues(self.theta(t))

    def gyro(self, t, h=1e-6):
        Rm, Rp = self.R(t - h), self.R(t + h)
        dR = Rm.T @ Rp
        return np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]]) / (4 * h)

    def accel(self, t):
        return self.R(t).T @ (self.acc(t) - G_S)


def periodic_trajectory(period=4.0, amp_scale=1.0):
    """A Trajectory that is exactly periodic with `period` seconds and passes through rest (p = v = theta = omega = 0) at t = k * period:
    every frequency is a multiple of 1 / period (the attitude terms A sin^2(2 pi f t) have period 1 / (2 f)).  One rendered period can
    then be replayed for ever as a physically consistent stream (bench.py)."""
    k = 1.0 / period
    return Trajectory(amp=np.array([0.6, 0.4, 0.2]) * amp_scale, freq=np.array([2 * k, 1 * k, 1 * k]),
                      rot_amp=np.array([0.10, 0.08, 0.25]) * amp_scale, rot_freq=np.array([1 * k, 1 * k, 0.5 * k]))


def make_world(n=1000, seed=0, xlim=(-4, 4), ylim=(1.0, 6.0), zlim=(-3, 3)):
    rng = np.random.default_rng(seed)
    return np.column_stack([rng.uniform(*xlim, n), rng.uniform(*ylim, n), rng.uniform(*zlim, n)])


def camera_pose(traj, t, Rbc, Tbc):
    Rsb, Tsb = traj.R(t), traj.pos(t)
    return Rsb @ Rbc, Rsb @ Tbc + Tsb


def observe_points(world, Rsc, Tsc, K, rows, cols, rng, sigma=0.5, zmin=0.3, zmax=4.5, border=10, cam=None):
    """Projects the world points visible from (Rsc, Tsc): pinhole (common/camera_pinhole.h:17-37) or, with cam["model"] ==
    "equidistant", the Kannala-Brandt model of common/camera_equidist.h:23-60 (r = theta (1 + k0 theta^2 + ...))."""
    Xc = (world - Tsc) @ Rsc
    z = Xc[:, 2]
    ok = (z > zmin) & (z < zmax)
    zs = np.where(ok, z, 1)
    xn, yn = Xc[:, 0] / zs, Xc[:, 1] / zs
    if cam is not None and cam.get("model", "pinhole") == "equidistant":
        k0, k1, k2, k3 = cam.get("k0123", (0, 0, 0, 0))
        th = np.arctan2(np.hypot(xn, yn), 1.0)
        phi = np.arctan2(yn, xn)
        t2 = th * th
        r = th * (1 + t2 * (k0 + t2 * (k1 + t2 * (k2 + t2 * k3))))
        u, v = K[0] * r * np.cos(phi) + K[2], K[1] * r * np.sin(phi) + K[3]
    else:
        u, v = K[0] * xn + K[2], K[1] * yn + K[3]
    ok &= (u > border) & (u < cols - border) & (v > border) & (v < rows - border)
    idx = np.nonzero(ok)[0]
    xp = np.column_stack([u[idx], v[idx]]) + rng.normal(0, sigma, (len(idx), 2))
    return idx.astype(np.int32), np.column_stack([xp, z[idx]])


def pcw_stream(cfg: dict, duration=4.0, imu_dt=0.005, vision_dt=0.04, seed=0, noise_accel=1e-4, noise_gyro=1e-5,
               pixel_sigma=0.5, npts=1500, traj=None):
    """Message list [(kind, ts_ns, payload)] in arrival order (IMU first on ties), like
    scripts/pyxivo_pcw.py:103-113.  kind 'imu' -> (gyro, accel); 'pc' -> (ids, xp_depth)."""
    traj = traj or Trajectory()
    rng = np.random.default_rng(seed + 1)
    cam = cfg["camera_cfg"]
    K = (cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    X = cfg["X"]
    Wbc = np.array(X["Wbc"], dtype=float)
    Rbc = rodrigues(Wbc) if Wbc.size == 3 else Wbc.reshape(3, 3)
    Tbc = np.array(X["Tbc"], dtype=float)
    world = make_world(npts, seed)
    msgs = []
    for t in np.arange(0, duration, imu_dt):
        msgs.append((t, 0, "imu", (traj.gyro(t) + rng.normal(0, noise_gyro, 3), traj.accel(t) + rng.normal(0, noise_accel, 3))))
    for t in np.arange(0, duration, vision_dt):
        Rsc, Tsc = camera_pose(traj, t, Rbc, Tbc)
        ids, xpd = observe_points(world, Rsc, Tsc, K, cam["rows"], cam["cols"], rng, pixel_sigma, cam=cam)
        so = cfg.get("sim_outliers")  # {"fraction", "pixels", "from_time"}: gross measurement errors for the outlier-rejection paths
        if so and t >= so.get("from_time", 1.0) and len(ids):
            ro = np.random.default_rng(seed * 7919 + int(round(t * 1000)))
            bad = ro.uniform(size=len(ids)) < so.get("fraction", 0.05)
            ang = ro.uniform(0, 2 * np.pi, len(ids))
            xpd = np.array(xpd, dtype=float, copy=True)
            xpd[bad, 0] += so.get("pixels", 3.0) * np.cos(ang[bad])
            xpd[bad, 1] += so.get("pixels", 3.0) * np.sin(ang[bad])
        msgs.append((t, 1, "pc", (ids, xpd)))
    msgs.sort(key=lambda m: (m[0], m[1]))
    return [(k, int(round(t * 1e9)), p) for (t, _, k, p) in msgs], traj


def pixel_rays(cam: dict):
    """Viewing ray (unit z for pinhole, unit norm for equidistant) and vignette weight of every pixel.
    pinhole: common/camera_pinhole.h:40-52; equidistant (Kannala-Brandt k0123): common/camera_equidist.h:97-160
    solved per pixel with Newton's method on theta."""
    rows, cols = cam["rows"], cam["cols"]
    v, u = np.mgrid[0:rows, 0:cols].astype(float)
    xn, yn = (u - cam["cx"]) / cam["fx"], (v - cam["cy"]) / cam["fy"]
    if cam.get("model", "pinhole") == "pinhole":
        return np.stack([xn, yn, np.ones_like(xn)], -1), None
    k0, k1, k2, k3 = cam.get("k0123", (0, 0, 0, 0))
    rd = np.hypot(xn, yn)
    th = rd.copy()
    for _ in range(12):
        t2 = th * th
        f = th * (1 + t2 * (k0 + t2 * (k1 + t2 * (k2 + t2 * k3)))) - rd
        df = 1 + t2 * (3 * k0 + t2 * (5 * k1 + t2 * (7 * k2 + t2 * 9 * k3)))
        th = th - f / df
    sc = np.sin(th) / np.maximum(rd, 1e-12)
    rays = np.stack([xn * sc, yn * sc, np.cos(th)], -1)
    # smooth vignette like a real fisheye: full brightness to 65 deg, dark beyond 80 deg (no hard edge for FAST)
    w = np.clip((np.deg2rad(80.0) - th) / np.deg2rad(15.0), 0.0, 1.0)
    return rays, w * w * (3 - 2 * w)


class PlaneRenderer:
    """Textured wall y = y0 seen by a pinhole or equidistant camera: each pixel ray is intersected with the
    plane and the texture is sampled bilinearly.  Physically consistent, so LK tracks are 3-D consistent."""

    def __init__(self, rows, cols, K, y0=3.0, seed=0, tex_scale=110.0, size=2048, cam=None):
        from . import synth

        self.rows, self.cols, self.K, self.y0, self.s, self.size = rows, cols, K, y0, tex_scale, size
        self.tex = synth.texture_canvas(size - 64, size - 64, seed, pad=32)
        cam = cam or dict(model="pinhole", rows=rows, cols=cols, fx=K[0], fy=K[1], cx=K[2], cy=K[3])
        self.rays, self.vignette = pixel_rays(cam)

    _Q = 1024  # sub-texel resolution of the sampling grid

    def _sample_fixed_point(self, tx, ty):
        """Bilinear sample of the texture in integer arithmetic: the sampling coordinates are snapped to a 1/1024-texel grid and the
        texture to 1/16 grey level, so the rendered frame does not depend on the host's libm / BLAS / SIMD rounding (a 1e-16 difference
        in a ray would need to straddle a grid point to change anything).  Mirror boundary like ndimage's mode="reflect"."""
        Q = self._Q
        if not hasattr(self, "_texq"):
            self._texq = np.rint(self.tex.astype(np.float64) * 16.0).astype(np.int64)
        R, C = self._texq.shape
        xq = np.rint(tx * Q).astype(np.int64)
        yq = np.rint(ty * Q).astype(np.int64)
        x0, fx = xq >> 10, xq & (Q - 1)
        y0, fy = yq >> 10, yq & (Q - 1)

        def refl(i, n):
            i = np.mod(i, 2 * n)
            return np.where(i >= n, 2 * n - 1 - i, i)

        xa, xb, ya, yb = refl(x0, C), refl(x0 + 1, C), refl(y0, R), refl(y0 + 1, R)
        t = self._texq
        val = (t[ya, xa] * (Q - fx) + t[ya, xb] * fx) * (Q - fy) + (t[yb, xa] * (Q - fx) + t[yb, xb] * fx) * fy
        return val.astype(np.float64) / (16.0 * Q * Q)  # < 2^32: exact in float64, division by a power of two is exact

    def render(self, Rsc, Tsc, noise_rng=None, fast=False):
        """fast=True uses cv2.remap (bench only: ~40x quicker, fixed-point bilinear weights)."""
        d = self.rays @ Rsc.T
        dy = d[..., 1] if self.vignette is None else np.maximum(d[..., 1], 0.05)  # wide-angle rays may miss the wall
        lam = (self.y0 - Tsc[1]) / dy
        X = Tsc[0] + lam * d[..., 0]
        Z = Tsc[2] + lam * d[..., 2]
        tx = X * self.s + self.size / 2
        ty = -Z * self.s + self.size / 2
        if fast:
            import cv2

            img = cv2.remap(self.tex, tx.astype(np.float32), ty.astype(np.float32), cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
        else:
            img = self._sample_fixed_point(tx, ty)
        if self.vignette is not None:
            img = img * self.vignette
        if noise_rng is not None:
            img = img + noise_rng.normal(0, 1.5, img.shape)
        return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def image_stream(cfg: dict, duration=2.0, imu_dt=0.005, vision_dt=0.04, seed=0, noise_accel=1e-4, noise_gyro=1e-5,
                 stationary=0.2, channels=1, traj=None, fast=False, rest_accel_is_gravity=False):
    """IMU + rendered frames.  `stationary` seconds of rest first so that gravity initialisation
    (estimator.cpp:439-473) sees still samples when simulation=false."""
    traj = traj or Trajectory()
    rng = np.random.default_rng(seed + 1)
    cam = cfg["camera_cfg"]
    K = (cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    X = cfg["X"]
    Wbc = np.array(X["Wbc"], dtype=float)
    Rbc = rodrigues(Wbc) if Wbc.size == 3 else Wbc.reshape(3, 3)
    Tbc = np.array(X["Tbc"], dtype=float)
    rend = PlaneRenderer(cam["rows"], cam["cols"], K, seed=seed, cam=cam)
    tt = lambda t: max(0.0, t - stationary)
    msgs = []
    for t in np.arange(0, duration, imu_dt):
        acc = traj.R(0.0).T @ (-G_S) if (rest_accel_is_gravity and t < stationary) else traj.accel(tt(t))  # bench: a platform at rest measures gravity only
        msgs.append((t, 0, "imu", (traj.gyro(tt(t)) * (t >= stationary) + rng.normal(0, noise_gyro, 3), acc + rng.normal(0, noise_accel, 3))))
    for t in np.arange(0, duration, vision_dt):
        Rsc, Tsc = camera_pose(traj, tt(t), Rbc, Tbc)
        img = rend.render(Rsc, Tsc, rng, fast)
        if channels == 3:
            img = np.repeat(img[:, :, None], 3, axis=2)
        msgs.append((t, 1, "img", np.ascontiguousarray(img)))
    msgs.sort(key=lambda m: (m[0], m[1]))
    return [(k, int(round(t * 1e9)), p) for (t, _, k, p) in msgs], traj
