"""ctypes front-end for the two CPU checkers (oracle/oracle_api.h).

kind="ref"  -> oracle/_ref/libwembed_ref.so : the reference's own C++ sources compiled in place
               (oracle/Makefile `ref`), third-party deps shimmed.  Authoritative.
kind="port" -> oracle/libwembed_port.so     : oracle/wembed_port.cpp, our CPU restatement.

TEST INFRASTRUCTURE - never imported by the product.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATHS = {"ref": os.path.join(_HERE, "_ref", "libwembed_ref.so"), "port": os.path.join(_HERE, "libwembed_port.so")}
_LIBS: dict = {}


class OrcOptions(C.Structure):
    _fields_ = [(k, C.c_int32) for k in (
        "embeddingDimension", "weightType", "optimizerType", "maxIterations", "lrScheduleType", "warmupSteps",
        "lrAdaptPatience", "stopCriterion", "stopDisplacementPatience", "lossRateWindow", "stopLossPatience",
        "numThreads")] + [(k, C.c_double) for k in (
        "dimensionHint", "attractionScale", "repulsionScale", "centreScale", "edgeLength", "doublingFactor",
        "simpleOptMaxDisplacement", "learningRate", "lrCoolingFactor", "lrDecayFactor", "lrDecayThreshold",
        "lrGrowthFactor", "lrGrowthThreshold", "stopDisplacementTol", "lossSmoothingFactor", "stopLossTol")]


def build(kind: str = "port", reference: str = "/root/reference") -> bool:
    """Compile a checker.  "ref" needs the reference checkout; returns False when it is absent."""
    if kind == "ref" and not os.path.isdir(reference):
        return os.path.exists(_PATHS["ref"])
    subprocess.run(["make", "-s", "-j8", "-C", _HERE, kind, f"REF={reference}"], check=True)
    return os.path.exists(_PATHS[kind])


def have(kind: str) -> bool:
    return os.path.exists(_PATHS[kind])


def _lib(kind: str):
    if kind in _LIBS:
        return _LIBS[kind]
    if not have(kind):
        raise FileNotFoundError(f"{_PATHS[kind]} missing: run `make -C oracle {kind}`")
    lib = C.CDLL(_PATHS[kind])
    P, i32, i64, dp, ip = C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_int32)
    sig = {
        "options_default": (None, [C.POINTER(OrcOptions)]),
        "create": (P, [i32, i64, ip, ip, C.POINTER(OrcOptions), i32, i32]),
        "destroy": (None, [P]),
        "num_vertices": (i32, [P]),
        "num_directed_edges": (i64, [P]),
        "csr": (None, [P, ip, ip]),
        "are_neighbors": (i32, [P, i32, i32]),
        "set_coordinates": (None, [P, dp]),
        "set_weights": (None, [P, dp]),
        "get_coordinates": (None, [P, dp]),
        "get_weights": (None, [P, dp]),
        "get_forces": (None, [P, dp]),
        "step": (None, [P]),
        "is_finished": (i32, [P]),
        "run": (i64, [P]),
        "get_stats": (None, [P, dp]),
        "candidates": (i64, [P, i32, ip, i64]),
    }
    for name, (res, args) in sig.items():
        f = getattr(lib, f"{kind}_{name}")
        f.restype, f.argtypes = res, args
    _LIBS[kind] = lib
    return lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class CpuEmbedder:
    """One WembedEmbedder (kind="ref") or its CPU restatement (kind="port")."""

    STATS = ("loss_attract", "loss_repel", "lr", "rel_displacement", "rel_loss_improvement", "iteration",
             "num_rep_pairs", "reserved")

    def __init__(self, kind, edges, n=None, seed=1234, init_state=True, **opts):
        self.kind, self._l = kind, _lib(kind)
        self.opts = OrcOptions()
        self._f("options_default")(C.byref(self.opts))
        for k, v in opts.items():
            if not hasattr(self.opts, k):
                raise AttributeError(k)
            setattr(self.opts, k, v)
        e = np.ascontiguousarray(np.asarray(edges, dtype=np.int32).reshape(-1, 2))
        src, dst = np.ascontiguousarray(e[:, 0]), np.ascontiguousarray(e[:, 1])
        n_hint = int(n) if n is not None else (int(e.max()) + 1 if e.size else 0)
        self._h = self._f("create")(n_hint, len(src), _ip(src), _ip(dst), C.byref(self.opts), seed, int(init_state))
        self.n = self._f("num_vertices")(self._h)
        self.d = self.opts.embeddingDimension

    def _f(self, name):
        return getattr(self._l, f"{self.kind}_{name}")

    def close(self):
        if getattr(self, "_h", None):
            self._f("destroy")(self._h)
            self._h = None

    __del__ = close

    def csr(self):
        rp = np.empty(self.n + 1, np.int32)
        col = np.empty(max(1, self._f("num_directed_edges")(self._h)), np.int32)
        self._f("csr")(self._h, _ip(rp), _ip(col))
        return rp, col[: rp[-1]]

    def are_neighbors(self, v, u):
        return bool(self._f("are_neighbors")(self._h, v, u))

    def set_coordinates(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(self.n, self.d)
        self._f("set_coordinates")(self._h, _dp(x))

    def set_weights(self, w):
        w = np.ascontiguousarray(w, dtype=np.float64).reshape(self.n)
        self._f("set_weights")(self._h, _dp(w))

    def _get(self, name, shape):
        out = np.empty(shape, np.float64)
        self._f(name)(self._h, _dp(out))
        return out

    def coordinates(self):
        return self._get("get_coordinates", (self.n, self.d))

    def weights(self):
        return self._get("get_weights", (self.n,))

    def forces(self):
        return self._get("get_forces", (self.n, self.d))

    def step(self):
        self._f("step")(self._h)

    def is_finished(self):
        return bool(self._f("is_finished")(self._h))

    def run(self):
        return int(self._f("run")(self._h))

    def stats(self):
        s = self._get("get_stats", (8,))
        return dict(zip(self.STATS, s.tolist()))

    def near_threshold(self, tau=1e-5):
        """Port only: vertices owning a pair within tau*L of the hinge threshold at the current coordinates."""
        assert self.kind == "port"
        f = self._l.port_flag_near_threshold
        f.restype, f.argtypes = None, [C.c_void_p, C.c_double, C.POINTER(C.c_uint8)]
        out = np.zeros(self.n, np.uint8)
        f(self._h, float(tau), out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out.astype(bool)

    def candidates(self, v):
        cap = 1024
        while True:
            out = np.empty(cap, np.int32)
            c = self._f("candidates")(self._h, int(v), _ip(out), cap)
            if c <= cap:
                return out[:c].copy()
            cap = int(c)


def ref_read_edge_list(path, comment="#", delimiter=" "):
    """The reference's GraphIO::readEdgeList + Graph construction on a file; returns (row_ptr, col) of the graph it built."""
    lib = _lib("ref")
    lib.ref_read_edge_list.restype = C.c_int32
    lib.ref_read_edge_list.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_int64)]
    lib.ref_graph_csr.restype = None
    lib.ref_graph_csr.argtypes = [C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    directed = C.c_int64()
    n = lib.ref_read_edge_list(str(path).encode(), comment.encode(), delimiter.encode(), C.byref(directed))
    rp, col = np.empty(n + 1, np.int32), np.empty(max(1, directed.value), np.int32)
    lib.ref_graph_csr(_ip(rp), _ip(col))
    return rp, col[: directed.value]
