"""csrc/workpool.h under concurrent drivers and pollers (tests/cpp/workpool_stress.cpp): every parallel-for item runs exactly once, an item
that throws resurfaces on the job's owner, the availability counter returns to zero.  Host code only."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pool_runs_every_item_once_and_survives_throwing_items(tmp_path):
    exe = str(tmp_path / "workpool_stress")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", os.path.join(ROOT, "tests", "cpp", "workpool_stress.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr
