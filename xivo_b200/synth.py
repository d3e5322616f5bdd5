"""Synthetic inputs shared by tests and bench.py (SURVEY.md §8d recipes).  numpy/scipy only."""
from __future__ import annotations

import numpy as np
from scipy import ndimage


def texture_canvas(rows: int, cols: int, seed: int = 0, pad: int = 32) -> np.ndarray:
    """Band-limited random texture (coarse uniform grid, cubic upsampling) + random constant
    squares, float32 canvas of (rows+2*pad) x (cols+2*pad)."""
    rng = np.random.default_rng(seed)
    R, Cc = rows + 2 * pad, cols + 2 * pad
    coarse = rng.uniform(0, 255, (R // 8 + 2, Cc // 8 + 2)).astype(np.float32)
    canvas = ndimage.zoom(coarse, 8, order=3)[:R, :Cc].copy()
    nsq = int(400 * (rows * cols) / (512 * 512))
    for _ in range(nsq):
        s = int(rng.integers(4, 13))
        y = int(rng.integers(0, R - s))
        x = int(rng.integers(0, Cc - s))
        canvas[y : y + s, x : x + s] = rng.uniform(0, 255)
    return canvas


def frame_from_canvas(canvas: np.ndarray, rows: int, cols: int, shift=(0, 0), noise_seed: int = 1, pad: int = 32) -> np.ndarray:
    dx, dy = shift
    img = canvas[pad + dy : pad + dy + rows, pad + dx : pad + dx + cols]
    rng = np.random.default_rng(noise_seed)
    img = img + rng.normal(0, 2, img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def frame_pair(rows: int, cols: int, seed: int = 0, shift=(3, 2)):
    c = texture_canvas(rows, cols, seed)
    return frame_from_canvas(c, rows, cols, (0, 0), seed * 2 + 1), frame_from_canvas(c, rows, cols, shift, seed * 2 + 2)


def to_bgr(gray: np.ndarray, distinct: bool = False) -> np.ndarray:
    """Grey -> 3-channel (what cv::imread hands the reference).  distinct=True decorrelates the
    channels a little so channel handling bugs cannot hide."""
    out = np.repeat(gray[:, :, None], 3, axis=2).copy()
    if distinct:
        out[..., 1] = np.roll(gray, 1, axis=0)
        out[..., 2] = np.roll(gray, 2, axis=1)
    return out


def random_rotation(rng, scale=1.0):
    w = rng.normal(0, scale, 3)
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3)
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th**2 * K @ K


def random_filter_problem(G: int, F: int, n: int, seed: int = 0, model: int = 0):
    """A random but physically plausible in-state configuration: camera, motion state, groups,
    features that project inside the image, covariance with the reference's structure
    (zero rows/cols for empty slots, SPD on live ones)."""
    rng = np.random.default_rng(seed)
    N = 23 + 6 * G + 3 * F
    if model == 0:
        camera = np.array([0, 480, 640, 275.0, 275.0, 320.0, 240.0, 0, 0, 0, 0], dtype=np.float64)
    else:
        camera = np.array([3, 512, 512, 190.98, 190.97, 254.93, 256.90, 0.0034, 0.0007, -0.0020, 0.0002], dtype=np.float64)
    Rsb = random_rotation(rng, 0.3)
    Tsb = rng.normal(0, 0.5, 3)
    Rbc = random_rotation(rng, 0.1) @ np.array([[1.0, 0, 0], [0, 0, 1], [0, -1, 0]])
    Tbc = rng.normal(0, 0.05, 3)
    X24 = np.concatenate([Rsb.ravel(), Tsb, Rbc.ravel(), Tbc])
    groups = np.zeros((G, 12))
    ng = min(G, max(2, n // 3))
    for g in range(G):
        Rg = Rsb @ random_rotation(rng, 0.05)
        Tg = Tsb + rng.normal(0, 0.1, 3)
        groups[g] = np.concatenate([Rg.ravel(), Tg])
    feat_x = np.zeros((n, 3))
    feat_xp = np.zeros((n, 2))
    feat_ref = rng.integers(0, ng, n).astype(np.int32)
    feat_sind = rng.permutation(F)[:n].astype(np.int32)
    for i in range(n):
        feat_x[i] = [rng.uniform(-0.6, 0.6), rng.uniform(-0.45, 0.45), np.log(rng.uniform(0.8, 4.0))]
        feat_xp[i] = [rng.uniform(40, camera[2] - 40), rng.uniform(40, camera[1] - 40)]
    # covariance: SPD on live dofs, zero elsewhere
    live = list(range(23))
    for g in range(ng):
        live += list(range(23 + 6 * g, 23 + 6 * g + 6))
    for s in feat_sind:
        live += list(range(23 + 6 * G + 3 * s, 23 + 6 * G + 3 * s + 3))
    live = np.array(sorted(live))
    A = rng.normal(0, 1, (len(live), len(live)))
    scale = np.exp(rng.uniform(-6, 0, len(live)))
    Pl = (A @ A.T / len(live) + np.eye(len(live))) * np.outer(scale, scale)
    P = np.zeros((N, N))
    P[np.ix_(live, live)] = Pl
    P = 0.5 * (P + P.T)
    return dict(G=G, F=F, N=N, camera=camera, X24=X24, groups=groups, feat_x=feat_x, feat_xp=feat_xp, feat_ref=feat_ref,
                feat_sind=feat_sind, P=P, n=n)


def set_measurements_near_prediction(prob, oracle_mod, sigma=1.0, seed=0):
    """Replace feat_xp by predicted pixel + noise so innovations look like a running filter."""
    rng = np.random.default_rng(seed)
    lay = oracle_mod.Layout(prob["G"], prob["F"])
    cam = camera_from_array(oracle_mod, prob["camera"])
    X = prob["X24"]
    Rsb, Tsb, Rbc, Tbc = X[:9].reshape(3, 3), X[9:12], X[12:21].reshape(3, 3), X[21:24]
    for i in range(prob["n"]):
        g = prob["groups"][prob["feat_ref"][i]]
        _, inn, cache = oracle_mod.feature_jacobian(lay, cam, Rsb, Tsb, Rbc, Tbc, g[:9].reshape(3, 3), g[9:12], prob["feat_x"][i],
                                                    np.zeros(2), int(prob["feat_ref"][i]), int(prob["feat_sind"][i]))
        prob["feat_xp"][i] = cache["xp"] + rng.normal(0, sigma, 2)
    return prob


def camera_from_array(oracle_mod, c):
    return oracle_mod.Camera(int(c[0]), int(c[1]), int(c[2]), c[3], c[4], c[5], c[6], tuple(c[7:11]))
