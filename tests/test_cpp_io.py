"""include/xivo_b200_io.hpp + examples/vio.cpp (the reference's DataLoader / `vio` app surface, src/loader.cpp,
src/app/vio.cpp): the C++ loader and PNM reader agree with xivo_b200.dataio on the same ASL folder (CPU); on a GPU box
the app's trajectory file equals what the Python API produces for the same messages."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from xivo_b200 import dataio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "xivo_b200")
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")


def build(tmp_path):
    exe = str(tmp_path / "vio")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "vio.cpp"), "-o", exe,
           "-L", LIBDIR, "-lxivo_b200", "-Wl,-rpath," + LIBDIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def make_folder(tmp_path, n=48, shape=(20, 24)):
    rng = np.random.default_rng(5)
    msgs = []
    for k in range(n):
        msgs.append(("imu", 1_403_636_579_000_000_000 + k * 5_000_000, (rng.normal(size=3), rng.normal(size=3) * 9.8)))
        if k % 8 == 0:
            shp = shape if (k // 8) % 2 == 0 else shape + (3,)
            msgs.append(("img", 1_403_636_579_000_000_000 + k * 5_000_000, rng.integers(0, 256, shp, dtype=np.uint8)))
    return msgs, dataio.write_asl(str(tmp_path / "seq"), msgs)


def test_cpp_loader_matches_python_loader(tmp_path):
    exe = build(tmp_path)
    msgs, (cam, imu) = make_folder(tmp_path)
    r = subprocess.run([exe, "--list", cam, imu], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    py = dataio.load_asl(cam, imu)
    assert len(lines) == len(py) == len(msgs)
    for line, (kind, ts, p) in zip(lines, py):  # same order (images first on equal stamps), same values
        tok = line.split()
        assert tok[0] == kind and int(tok[1]) == ts
        if kind == "imu":
            assert np.array_equal(np.array([float(x) for x in tok[2:8]]), np.concatenate(p))
        else:
            img = dataio.read_pnm(p)
            shape = "x".join(str(s) for s in (img.shape if img.ndim == 3 else img.shape + (1,)))
            assert tok[2] == shape and int(tok[3]) == int(img.astype(np.uint64).sum())
    # missing folder -> error exit, like the reference's LOG(FATAL)
    assert subprocess.run([exe, "--list", str(tmp_path / "nope"), imu], capture_output=True).returncode == 2


def test_cpp_pnm_reader_returns_bgr(tmp_path):
    exe = build(tmp_path)
    img = np.zeros((4, 5, 3), np.uint8)
    img[..., 0] = 7  # B plane only
    cam, imu = dataio.write_asl(str(tmp_path / "s"), [("img", 10, img), ("imu", 10, (np.zeros(3), np.zeros(3)))])
    assert np.array_equal(dataio.read_pnm(os.path.join(cam, "data", "10.ppm")), img)
    out = subprocess.run([exe, "--list", cam, imu], capture_output=True, text=True).stdout.split()
    assert out[:2] == ["img", "10"] and int(out[3]) == 7 * 20  # image first on the tie, checksum of the single non-zero plane


@pytest.mark.gpu
def test_vio_app_writes_the_same_trajectory_as_the_python_api(tmp_path):
    from xivo_b200 import pyxivo, sim

    exe = build(tmp_path)
    cfg_path = os.path.join(LIBDIR, "cfg", "vio_640x480.json")
    cfg = sim.load_cfg(cfg_path)
    cfg["camera_cfg"].update(rows=240, cols=320, fx=137.5, fy=137.5, cx=160, cy=120)
    cfg["tracker_cfg"].update(num_features_min=60, num_features_max=80)
    import json

    small = str(tmp_path / "cfg.json")
    json.dump(cfg, open(small, "w"))
    msgs, _ = sim.image_stream(cfg, duration=1.2, seed=6)
    cam, imu = dataio.write_asl(str(tmp_path / "seq"), msgs)
    out = str(tmp_path / "out_state")
    r = subprocess.run([exe, small, cam, imu, out, "4", "14"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    stamps, poses = dataio.read_vio_trajectory(out)
    b = pyxivo.Batch(cfg, n_seq=1, max_groups=4, max_features=14)
    ref_ts, ref_g = [], []
    for kind, ts, p in dataio.load_asl(cam, imu):
        if kind == "imu":
            b.inertial_meas(ts, p[0], p[1])
        else:
            b.visual_meas(ts, [dataio.read_pnm(p)])
        ref_ts.append(b.now(0))
        ref_g.append(b.gsb(0))
    assert len(stamps) == len(ref_ts) and np.array_equal(stamps, np.array(ref_ts, dtype=np.int64))
    assert np.abs(poses - np.array(ref_g)).max() <= 1e-7  # 9 significant digits in the text file
    assert b.counters(0)["num_instate_features"] > 0
    b.close()
