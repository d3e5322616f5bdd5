"""The C-ABI library loads and exports every symbol include/*.h declares (no compute, no GPU)."""
import ctypes
import glob
import os
import re

from xivo_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    syms = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        for m in re.finditer(r"\b(xivo_[a-z0-9_]+)\s*\(", src):
            syms.add(m.group(1))
    return syms


def test_library_exports_all_declared_symbols():
    assert os.path.exists(capi.LIB_PATH), "build libxivo_b200.so first (python -m xivo_b200.build)"
    lib = ctypes.CDLL(capi.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 15
    missing = [s for s in sorted(syms) if not hasattr(lib, s)]
    assert not missing, f"declared in include/ but not exported: {missing}"


def test_no_cpu_fallback_without_device():
    """Without a CUDA device ctx creation must fail loudly (code XIVO_ERR_CUDA), never fall back."""
    lib = capi.lib()
    h = ctypes.c_void_p()
    rc = lib.xivo_ctx_create(0, ctypes.byref(h))
    if rc == 0:  # running on a GPU box: fine, just clean up
        lib.xivo_ctx_destroy(h)
    else:
        assert rc == -2
        assert b"no CPU path" in lib.xivo_last_error()


def test_pyramid_layout_matches_opencv_level_rule():
    total, lv = capi.Context.pyramid_layout(480, 640, 1, 15, 5)
    assert [(r, c) for r, c, _ in lv] == [(480, 640), (240, 320), (120, 160), (60, 80), (30, 40)]
    total, lv = capi.Context.pyramid_layout(512, 512, 3, 15, 5)
    assert len(lv) == 6 and lv[-1][:2] == (16, 16) and total >= sum(r * c * 3 for r, c, _ in lv)


def test_homography_mask_through_the_c_abi_matches_the_oracle():
    """xivo_find_homography_mask is host work: the product library itself can be checked without a GPU against the restatement of
    cv::findHomography (oracle/homography_oracle.py, pinned on cv2 4.13)."""
    import numpy as np

    from oracle import homography_oracle as HO

    lib = capi.lib()
    rng = np.random.default_rng(3)
    for method in (HO.LMEDS, HO.RANSAC):
        for _ in range(6):
            n = int(rng.integers(20, 150))
            p0 = rng.uniform(8, 630, (n, 2)).astype(np.float32)
            p1 = (p0 + np.float32([4, -2]) + rng.normal(0, 0.5, (n, 2))).astype(np.float32)
            p1[: n // 5] += rng.normal(0, 25, (n // 5, 2)).astype(np.float32)
            mask, ok = np.zeros(n, np.uint8), ctypes.c_int()
            rc = lib.xivo_find_homography_mask(p0.ctypes.data_as(ctypes.c_void_p), p1.ctypes.data_as(ctypes.c_void_p), n, method, ctypes.c_double(3.0), 2000,
                                               ctypes.c_double(0.995), mask.ctypes.data_as(ctypes.c_void_p), ctypes.byref(ok))
            assert rc == 0 and ok.value == 1
            want_ok, want, _H = HO.find_homography_mask_413(p0, p1, method, 3.0, 2000, 0.995)
            assert want_ok and int((mask != want).sum()) <= 1 and n // 5 <= int((mask == 0).sum()) <= n // 5 + 3
    assert lib.xivo_find_homography_mask(None, None, 0, 4, ctypes.c_double(3.0), 10, ctypes.c_double(0.9), None, None) == -1


def test_new_getters_reject_a_null_batch_without_touching_a_device():
    """Every read-back entry point added for the binding's surface returns XIVO_ERR_ARG (-1) for a null handle (no throw, no crash)."""
    lib = capi.lib()
    n = ctypes.c_int()
    null = ctypes.c_void_p()
    assert lib.xivo_get_instate_feature_table(null, 0, -1, None, None, None, None, None, None, None, None, None, 0, ctypes.byref(n)) == -1
    assert lib.xivo_get_instate_group_table(null, 0, None, None, None, None, 0, ctypes.byref(n)) == -1
    assert lib.xivo_get_calibration(null, 0, None, None, None, None, None) == -1
    assert lib.xivo_get_just_dropped(null, 0, None, 0, ctypes.byref(n)) == -1
    assert lib.xivo_get_tracker_counters(null, 0, (ctypes.c_int * 4)()) == -1
    assert lib.xivo_scale_init_velocity(null, 0, ctypes.c_double(2.0)) == -1
    assert b"null batch" in lib.xivo_last_error()
    prev = lib.xivo_set_frame_ingest(1)
    assert prev in (0, 1) and lib.xivo_set_frame_ingest(7) == 1 and lib.xivo_set_frame_ingest(prev) == 1  # an unknown mode only queries
