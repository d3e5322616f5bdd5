"""include/xivo_b200.hpp (C++ facade with the reference's Estimator / Tracker method names) compiles against the
C ABI and links with the library; without a GPU it must fail with XIVO_ERR_CUDA, on a GPU box it runs a smoke."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "facade_check.cpp")
LIBDIR = os.path.join(ROOT, "xivo_b200")


def build(tmp_path):
    exe = str(tmp_path / "facade_check")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe, "-L", LIBDIR, "-lxivo_b200",
           "-Wl,-rpath," + LIBDIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_facade_compiles_links_and_has_no_cpu_fallback(tmp_path):
    exe = build(tmp_path)
    r = subprocess.run([exe, os.path.join(LIBDIR, "cfg", "pcw_sim.json")], capture_output=True, text=True)
    # 42 = xivo::Error with XIVO_ERR_CUDA (no device here); 0 = ran on a GPU
    assert r.returncode in (0, 42), r.stdout + r.stderr
    if r.returncode == 42:
        assert "no CPU path" in r.stdout


@pytest.mark.gpu
def test_facade_smoke_on_gpu(tmp_path):
    exe = build(tmp_path)
    r = subprocess.run([exe, os.path.join(LIBDIR, "cfg", "pcw_sim.json")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok N=89" in r.stdout and "tracker ok" in r.stdout
