"""ctypes binding of the C ABI in include/wembed_b200.h (libwembed_b200.so).

This is the thinnest possible Python view of the drop-in boundary: every method is one C call.
The library is built in-tree by `python -m wembed_b200.build` (or __graft_entry__.build()).
There is no fallback: if the shared library is missing, importing fails loudly; if there is no
CUDA device, wb_create fails with WB_ERR_NO_DEVICE.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WB_LIB", os.path.join(_HERE, "lib", "libwembed_b200.so"))   # WB_LIB: A/B builds of the same ABI

WB_OK, WB_ERR_INVALID, WB_ERR_CUDA, WB_ERR_NO_DEVICE, WB_ERR_UNSUPPORTED = 0, -1, -2, -3, -4
WB_OPT_SIMPLE, WB_OPT_ADAM = 0, 1

# every symbol include/wembed_b200.h declares
EXPORTS = (
    "wb_abi_version", "wb_build_info", "wb_last_error", "wb_device_count", "wb_options_default", "wb_create", "wb_destroy",
    "wb_set_coordinates", "wb_set_weights", "wb_get_coordinates", "wb_get_weights", "wb_get_forces", "wb_reset_optimizer",
    "wb_set_iteration", "wb_step", "wb_step_async", "wb_step_collect", "wb_synchronize", "wb_query_candidates",
    "wb_enable_timing", "wb_get_phase_times", "wb_mark", "wb_elapsed_ms", "wb_launch_count", "wb_comm_unique_id", "wb_comm_init", "wb_reconstruction",
    "wb_edge_detection", "wb_set_list_policy", "wb_get_partition", "wb_exec_mode", "wb_comm_init_local", "wb_step_group",
)


class WbOptions(C.Structure):
    _fields_ = [
        ("embedding_dimension", C.c_int32), ("optimizer", C.c_int32), ("reserved2", C.c_int32), ("device", C.c_int32),
        ("keep_forces", C.c_int32), ("reserved0", C.c_int32),
        ("attraction_scale", C.c_double), ("repulsion_scale", C.c_double), ("centre_scale", C.c_double),
        ("edge_length", C.c_double), ("doubling_factor", C.c_double), ("simple_max_displacement", C.c_double),
        ("seed", C.c_uint32), ("reserved1", C.c_uint32 * 7),
    ]


class WbStepStats(C.Structure):
    _fields_ = [
        ("loss_attract", C.c_double), ("loss_repel", C.c_double), ("sum_displacement", C.c_double),
        ("sum_radius_sq", C.c_double), ("rel_displacement", C.c_double), ("num_repulsion_pairs", C.c_double),
        ("num_candidates", C.c_double), ("num_box_tests", C.c_double), ("centroid", C.c_double * 32), ("iteration", C.c_int64),
        ("num_listed_pairs", C.c_double), ("list_rebuilt", C.c_double), ("list_skin", C.c_double), ("max_displacement_ratio", C.c_double),
    ]

    def as_dict(self):
        return {k: (list(getattr(self, k)) if k == "centroid" else getattr(self, k)) for k, _ in self._fields_}


class WbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"wembed_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing - build it with `python -m wembed_b200.build` (no CPU fallback exists)")
    l = C.CDLL(LIB_PATH)
    H, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    sig = {
        "wb_abi_version": (C.c_int, []), "wb_build_info": (C.c_char_p, []), "wb_last_error": (C.c_char_p, []),
        "wb_device_count": (C.c_int, []), "wb_options_default": (None, [C.POINTER(WbOptions)]),
        "wb_create": (C.c_int, [C.POINTER(H), i32, ip, ip, C.POINTER(WbOptions)]), "wb_destroy": (C.c_int, [H]),
        "wb_set_coordinates": (C.c_int, [H, dp]), "wb_set_weights": (C.c_int, [H, dp]),
        "wb_get_coordinates": (C.c_int, [H, dp]), "wb_get_weights": (C.c_int, [H, dp]), "wb_get_forces": (C.c_int, [H, dp]),
        "wb_reset_optimizer": (C.c_int, [H]), "wb_set_iteration": (C.c_int, [H, i64]),
        "wb_step": (C.c_int, [H, C.c_double, C.POINTER(WbStepStats)]), "wb_step_async": (C.c_int, [H, C.c_double]),
        "wb_step_collect": (C.c_int, [H, C.POINTER(WbStepStats)]), "wb_synchronize": (C.c_int, [H]),
        "wb_query_candidates": (C.c_int, [H, i32, ip, lp, ip, i64]),
        "wb_enable_timing": (C.c_int, [H, C.c_int]), "wb_get_phase_times": (C.c_int, [H, dp]),
        "wb_mark": (C.c_int, [H, C.c_int]), "wb_elapsed_ms": (C.c_int, [H, C.c_int, C.c_int, dp]),
        "wb_launch_count": (C.c_int64, [H]),
        "wb_reconstruction": (C.c_int, [H, i32, ip, dp]),
        "wb_edge_detection": (C.c_int, [H, C.c_int64, ip, ip, C.POINTER(C.c_uint8), dp]),
        "wb_set_list_policy": (C.c_int, [H, C.c_double, C.c_double]),
        "wb_get_partition": (C.c_int, [H, ip, ip]),
        "wb_exec_mode": (C.c_int, [H, C.c_char_p, i32]),
        "wb_comm_init_local": (C.c_int, [C.POINTER(H), i32]),
        "wb_step_group": (C.c_int, [C.POINTER(H), i32, C.c_double, C.POINTER(WbStepStats)]),
        "wb_comm_unique_id": (C.c_int, [C.c_char_p]), "wb_comm_init": (C.c_int, [H, C.c_char_p, i32, i32]),
    }
    for name, (res, args) in sig.items():
        f = getattr(l, name)
        f.restype, f.argtypes = res, args
    _lib = l
    return l


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = lib().wb_comm_unique_id(buf)
    if rc != WB_OK:
        raise WbError(rc, lib().wb_last_error().decode())
    return buf.raw


def default_options(**kw) -> WbOptions:
    o = WbOptions()
    lib().wb_options_default(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class DeviceEmbedder:
    """One wb_embedder handle."""

    def __init__(self, row_ptr, col, **opts):
        self._l = lib()
        self.opts = default_options(**opts)
        self.row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
        self.col = np.ascontiguousarray(col, dtype=np.int32)
        self.n = len(self.row_ptr) - 1
        self.d = self.opts.embedding_dimension
        self._h = C.c_void_p()
        colp = self.col if len(self.col) else np.zeros(1, np.int32)
        self._check(self._l.wb_create(C.byref(self._h), self.n, _ip(self.row_ptr), _ip(colp), C.byref(self.opts)))

    def _check(self, rc):
        if rc != WB_OK:
            raise WbError(rc, self._l.wb_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            self._l.wb_destroy(self._h)
            self._h = None

    __del__ = close

    def set_coordinates(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(self.n, self.d)
        self._check(self._l.wb_set_coordinates(self._h, _dp(x)))

    def set_weights(self, w):
        w = np.ascontiguousarray(w, dtype=np.float64).reshape(self.n)
        self._check(self._l.wb_set_weights(self._h, _dp(w)))

    def coordinates(self):
        out = np.empty((self.n, self.d), np.float64)
        self._check(self._l.wb_get_coordinates(self._h, _dp(out)))
        return out

    def weights(self):
        out = np.empty(self.n, np.float64)
        self._check(self._l.wb_get_weights(self._h, _dp(out)))
        return out

    def forces(self):
        out = np.empty((self.n, self.d), np.float64)
        self._check(self._l.wb_get_forces(self._h, _dp(out)))
        return out

    def reset_optimizer(self):
        self._check(self._l.wb_reset_optimizer(self._h))

    def set_iteration(self, it):
        self._check(self._l.wb_set_iteration(self._h, int(it)))

    def step(self, lr):
        st = WbStepStats()
        self._check(self._l.wb_step(self._h, float(lr), C.byref(st)))
        return st.as_dict()

    def step_async(self, lr):
        self._check(self._l.wb_step_async(self._h, float(lr)))

    def step_collect(self):
        st = WbStepStats()
        self._check(self._l.wb_step_collect(self._h, C.byref(st)))
        return st.as_dict()

    def set_list_policy(self, skin_max, reuse_steps=4.0):
        """Policy of the repulsion pair list (include/wembed_b200.h); results never depend on it."""
        self._check(self._l.wb_set_list_policy(self._h, float(skin_max), float(reuse_steps)))

    def synchronize(self):
        self._check(self._l.wb_synchronize(self._h))

    def enable_timing(self, on=True):
        self._check(self._l.wb_enable_timing(self._h, int(on)))

    def phase_times(self):
        out = np.empty(6, np.float64)
        self._check(self._l.wb_get_phase_times(self._h, _dp(out)))
        return dict(zip(("index", "attract_update", "repel", "unused", "recentre_observe", "total"), out.tolist()))

    def mark(self, slot):
        self._check(self._l.wb_mark(self._h, int(slot)))

    def elapsed_ms(self, a, b):
        out = C.c_double()
        self._check(self._l.wb_elapsed_ms(self._h, int(a), int(b), C.byref(out)))
        return out.value

    def launch_count(self):
        return int(self._l.wb_launch_count(self._h))

    def reconstruction(self, nodes):
        """(constructDeg, MAP) of the current layout for the sampled vertices (evaluationLib Reconstruction)."""
        q = np.ascontiguousarray(nodes, dtype=np.int32)
        out = np.zeros(2, np.float64)
        self._check(self._l.wb_reconstruction(self._h, len(q), _ip(q), _dp(out)))
        return float(out[0]), float(out[1])

    def edge_detection(self, v, w, is_edge):
        """(precision, recall, F1) at the best similarity threshold over the sampled pairs (evaluationLib EdgeDetection);
        wembed_b200.metrics.sample_edge_pairs mirrors the reference's sampler."""
        a = np.ascontiguousarray(v, dtype=np.int32)
        b = np.ascontiguousarray(w, dtype=np.int32)
        f = np.ascontiguousarray(is_edge, dtype=np.uint8)
        assert len(a) == len(b) == len(f)
        out = np.zeros(3, np.float64)
        self._check(self._l.wb_edge_detection(self._h, len(a), _ip(a), _ip(b), f.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(out)))
        return float(out[0]), float(out[1]), float(out[2])

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        """Join the vertex-sharded multi-GPU step (see include/wembed_b200.h)."""
        assert len(unique_id) == 128
        self._check(self._l.wb_comm_init(self._h, unique_id, int(rank), int(world)))

    def exec_mode(self):
        """("graph" | "direct", note): how steps are issued (include/wembed_b200.h: wb_exec_mode)."""
        buf = C.create_string_buffer(256)
        rc = self._l.wb_exec_mode(self._h, buf, 256)
        if rc < 0:
            self._check(rc)
        return ("graph" if rc == 1 else "direct"), buf.value.decode()

    def partition(self):
        """[begin, end) of the vertices this handle owns."""
        b, e = C.c_int32(), C.c_int32()
        self._check(self._l.wb_get_partition(self._h, C.byref(b), C.byref(e)))
        return b.value, e.value

    def query_candidates(self, queries):
        q = np.ascontiguousarray(queries, dtype=np.int32)
        offs = np.zeros(len(q) + 1, np.int64)
        cap = max(1024, 64 * len(q))
        while True:
            ids = np.empty(cap, np.int32)
            rc = self._l.wb_query_candidates(self._h, len(q), _ip(q), offs.ctypes.data_as(C.POINTER(C.c_int64)), _ip(ids), cap)
            if rc == WB_ERR_INVALID and offs[-1] > cap:
                cap = int(offs[-1])
                continue
            self._check(rc)
            return [ids[offs[i]:offs[i + 1]].copy() for i in range(len(q))]


def comm_init_local(devs):
    """Join DeviceEmbedder handles of this process (one device) into a sharded group; step them from one thread each."""
    arr = (C.c_void_p * len(devs))(*[d._h for d in devs])
    rc = lib().wb_comm_init_local(arr, len(devs))
    if rc != WB_OK:
        raise WbError(rc, lib().wb_last_error().decode())


def step_group(devs, lr):
    """One step of a local group (comm_init_local): wb_step_group; returns the handles' step records."""
    hs = (C.c_void_p * len(devs))(*[d._h for d in devs])
    st = (WbStepStats * len(devs))()
    rc = lib().wb_step_group(hs, len(devs), float(lr), st)
    if rc != WB_OK:
        raise WbError(rc, lib().wb_last_error().decode())
    return [st[i].as_dict() for i in range(len(devs))]


def csr_from_edges(n: int, edges) -> tuple[np.ndarray, np.ndarray]:
    """Sort-based CSR with the invariants of the reference's Graph (Graph.cpp:87-150): symmetric,
    deduplicated, rows ascending, self loops dropped."""
    e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    e = e[e[:, 0] != e[:, 1]]
    both = np.concatenate([e, e[:, ::-1]])
    key = both[:, 0] * n + both[:, 1]
    key.sort()                                   # (np.unique is two orders of magnitude slower than sort + mask in numpy 2.3)
    if len(key):
        key = key[np.concatenate(([True], key[1:] != key[:-1]))]
    src, dst = key // n, key % n
    row_ptr = np.zeros(n + 1, np.int64)
    row_ptr[1:] = np.cumsum(np.bincount(src, minlength=n))
    return row_ptr.astype(np.int32), dst.astype(np.int32)
