"""ctypes view of datagen.cpp (built in-tree with g++ on first use).  Benchmark / test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "datagen.cpp")
_OUT = os.path.join(os.path.dirname(_HERE), "lib", "libwembed_datagen.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_OUT) or os.path.getmtime(_SRC) > os.path.getmtime(_OUT):
            os.makedirs(os.path.dirname(_OUT), exist_ok=True)
            tmp = _OUT + f".tmp{os.getpid()}"
            subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-o", tmp, _SRC], check=True)
            os.replace(tmp, _OUT)
        l = C.CDLL(_OUT)
        l.wbd_pairs_within.restype = C.c_int64
        l.wbd_pairs_within.argtypes = [C.c_int64, C.c_void_p, C.c_double]
        l.wbd_take_edges.argtypes = [C.c_void_p]
        l.wbd_girg_pairs.restype = C.c_int64
        l.wbd_girg_pairs.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int]
        l.wbd_csr_canonical.restype = C.c_int
        l.wbd_csr_canonical.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = l
    return _lib


def pairs_within(points: np.ndarray, radius: float) -> np.ndarray:
    """All index pairs (i < j) with ||p_i - p_j|| < radius as int32 [m, 2], lexicographic order."""
    pts = np.ascontiguousarray(points, dtype=np.float64)
    m = lib().wbd_pairs_within(len(pts), pts.ctypes.data, float(radius))
    out = np.empty((m, 2), np.int32)
    lib().wbd_take_edges(out.ctypes.data)
    return out


def girg_pairs(points: np.ndarray, weights: np.ndarray, c: float, W: float, count_only: bool = False):
    """Threshold-GIRG edges (i < j, lexicographic) of datasets.heavy_tailed_graph, or only their number."""
    pts = np.ascontiguousarray(points, dtype=np.float64)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    m = lib().wbd_girg_pairs(len(pts), pts.ctypes.data, w.ctypes.data, float(c), float(W), int(count_only))
    if count_only:
        return int(m)
    out = np.empty((m, 2), np.int32)
    lib().wbd_take_edges(out.ctypes.data)
    return out


def csr_canonical(n: int, edges: np.ndarray):
    """CSR (row_ptr, col) of a unique, sorted edge list with src < dst; None if the list is not of that form."""
    e = np.ascontiguousarray(edges, dtype=np.int32).reshape(-1, 2)
    row_ptr, col = np.empty(n + 1, np.int32), np.empty(2 * len(e), np.int32)
    rc = lib().wbd_csr_canonical(n, len(e), e.ctypes.data, row_ptr.ctypes.data, col.ctypes.data)
    return (row_ptr, col) if rc == 0 else None
