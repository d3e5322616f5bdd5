"""Implementation-defined orders of the reference's C++ standard library, made available to the Python oracle.

Three decisions of the reference depend on libstdc++ rather than on the algorithm:
  * iteration of `std::unordered_map<int, FeaturePtr> features_` / `groups_` in GraphBase::GetFeaturesIf / GetGroupsIf
    (src/graphbase.cpp:124-146) — the gauge-feature candidates (src/graph.cpp:291), the group candidates and their features in
    AddGroupOfFeatures (src/manager.cpp:479-497), the unsorted "median" of AdaptInitialDepth (src/manager.cpp:257-266);
  * the unstable std::sort on candidates with tied keys (src/manager.cpp:375-376, :420-421, :492, :499-500).
The oracle mirrors the reference's insert / erase history into real std::unordered_map<int,int> containers and runs the real
std::sort (oracle/stdumap.cpp, compiled with the same g++ / libstdc++ as the reference build in oracle/_ref and as the product's
host code, which keeps the same containers).  With them the oracle agrees with the REFERENCE'S OWN ESTIMATOR (oracle/ref_runner.py)
to ~1e-14 on point-cloud streams (tests/test_reference_pin.py).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "_build", "libstdumap.so")
        src = os.path.join(HERE, "stdumap.cpp")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            os.makedirs(os.path.dirname(so), exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", src, "-o", so])
        L = C.CDLL(so)
        L.um_new.restype = C.c_void_p
        L.um_delete.argtypes = [C.c_void_p]
        L.um_insert.argtypes = [C.c_void_p, C.c_int]
        L.um_erase.argtypes = [C.c_void_p, C.c_int]
        L.um_keys.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.std_sort_candidates.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.std_sort_desc.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.mh_new.restype = C.c_void_p
        L.mh_new.argtypes = [C.c_int]
        L.mh_delete.argtypes = [C.c_void_p]
        L.mh_push.argtypes = [C.c_void_p, C.c_ulonglong, C.c_int, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


class StdUnorderedIntMap:
    """Keys of a std::unordered_map<int, T> in iteration order."""

    def __init__(self):
        self.h = _lib().um_new()

    def insert(self, k):
        _lib().um_insert(self.h, int(k))

    def erase(self, k):
        _lib().um_erase(self.h, int(k))

    def keys(self):
        buf = (C.c_int * 16384)()
        n = _lib().um_keys(self.h, buf, 16384)
        assert n <= 16384
        return [buf[i] for i in range(n)]

    def __del__(self):
        try:
            _lib().um_delete(self.h)
        except Exception:
            pass


def std_sort_candidates(feats):
    """std::sort(begin, end, Criteria::CandidateComparison) on `feats` in their current order."""
    n = len(feats)
    if n < 2:
        return list(feats)
    st = np.array([int(f.status) for f in feats], dtype=np.int32)
    sc = np.array([-f.P[2, 2] for f in feats], dtype=np.float64)  # Feature::score(), src/feature.cpp:133-141
    perm = np.zeros(n, dtype=np.int32)
    _lib().std_sort_candidates(n, st.ctypes.data, sc.ctypes.data, perm.ctypes.data)
    return [feats[i] for i in perm]


def std_sort_desc(items, counts):
    n = len(items)
    if n < 2:
        return list(items)
    c = np.array(counts, dtype=np.int32)
    perm = np.zeros(n, dtype=np.int32)
    _lib().std_sort_desc(n, c.ctypes.data, perm.ctypes.data)
    return [items[i] for i in perm]


class StdMessageHeap:
    """Estimator::MaintainBuffer (src/estimator.cpp:923-941) with the reference's timestamp-only comparator on libstdc++'s heap
    algorithms: push(ts, item) returns the item that becomes due (or None)."""

    def __init__(self, max_size=10):
        self.h = _lib().mh_new(int(max_size))
        self.items = {}
        self.next = 0

    def push(self, ts, item):
        k = self.next
        self.next += 1
        self.items[k] = item
        ots, oh = C.c_ulonglong(), C.c_int()
        if _lib().mh_push(self.h, C.c_ulonglong(int(ts)), k, C.byref(ots), C.byref(oh)):
            return self.items.pop(oh.value)
        return None

    def __del__(self):
        try:
            _lib().mh_delete(self.h)
        except Exception:
            pass
