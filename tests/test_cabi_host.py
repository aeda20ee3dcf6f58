"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol of include/wembed_b200.h,
and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from wembed_b200 import build, cabi
    build.build()
    return cabi


def test_header_symbols_exported(lib):
    header = open(os.path.join(ROOT, "include", "wembed_b200.h")).read()
    declared = set(re.findall(r"\b(wb_[a-z_0-9]+)\s*\(", header)) - {"wb_embedder"}
    assert declared == set(lib.EXPORTS), declared ^ set(lib.EXPORTS)
    l = lib.lib()
    for name in declared:
        assert getattr(l, name) is not None
    assert l.wb_abi_version() == 2
    assert b"sm_100a" in l.wb_build_info()


def test_options_default_mirror_embedder_options(lib):
    o = lib.default_options()
    assert (o.embedding_dimension, o.optimizer, o.attraction_scale, o.repulsion_scale, o.centre_scale, o.edge_length,
            o.doubling_factor, o.simple_max_displacement) == (4, 1, 1.0, 1.0, 0.0, 1.0, 2.0, 1.0)


def test_struct_layout_matches_header(lib, tmp_path):
    """sizeof / offsetof of the C structs as compiled by gcc == the ctypes mirror."""
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "wembed_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(wb_options),'
                   ' offsetof(wb_options, attraction_scale), offsetof(wb_options, seed), sizeof(wb_step_stats), offsetof(wb_step_stats, iteration));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(t) for t in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [C.sizeof(lib.WbOptions), lib.WbOptions.attraction_scale.offset, lib.WbOptions.seed.offset,
                   C.sizeof(lib.WbStepStats), lib.WbStepStats.iteration.offset]


def test_no_cpu_fallback(lib):
    if lib.lib().wb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(lib.WbError) as e:
        lib.DeviceEmbedder(np.array([0, 1, 2], np.int32), np.array([1, 0], np.int32))
    assert e.value.code == lib.WB_ERR_NO_DEVICE


def test_create_validates_csr(lib):
    l = lib.lib()
    h = C.c_void_p()
    o = lib.default_options()
    bad = [
        (np.array([0, 1, 2], np.int32), np.array([0, 0], np.int32)),      # self loop
        (np.array([0, 2, 2], np.int32), np.array([1, 1], np.int32)),      # not strictly ascending
        (np.array([0, 1, 2], np.int32), np.array([1, 5], np.int32)),      # out of range
        (np.array([1, 1, 2], np.int32), np.array([1, 0], np.int32)),      # row_ptr[0] != 0
        (np.array([0, 1, 1], np.int32), np.array([1], np.int32)),         # not symmetric: 0 -> 1 without 1 -> 0
    ]
    for rp, col in bad:
        rc = l.wb_create(C.byref(h), 2, rp.ctypes.data_as(C.POINTER(C.c_int32)), col.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(o))
        assert rc in (lib.WB_ERR_INVALID, lib.WB_ERR_NO_DEVICE)
        if l.wb_device_count() > 0:
            assert rc == lib.WB_ERR_INVALID
    o.embedding_dimension = 33
    rp, col = np.array([0, 1, 2], np.int32), np.array([1, 0], np.int32)
    assert l.wb_create(C.byref(h), 2, rp.ctypes.data_as(C.POINTER(C.c_int32)), col.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(o)) == lib.WB_ERR_UNSUPPORTED


def test_mt19937_restatement_matches_libstdcxx(tmp_path):
    """wembed_b200/csrc/mt19937.cuh (host+device) == std::seed_seq + std::mt19937 + std::normal_distribution."""
    src = tmp_path / "mt.cpp"
    src.write_text(r'''
#include <random>
#include <cstdio>
#include "mt19937.cuh"
int main() {
  int bad = 0;
  for (uint32_t seed : {1234u, 0u, 4294967295u}) for (uint32_t v = 0; v < 40; v++) for (uint32_t it : {1u, 7u, 100000u}) for (int d : {2, 3, 8}) {
    std::seed_seq seq{seed, v, it}; std::mt19937 gen(seq);
    double ref[8], norm = 0; for (int k = 0; k < d; k++) { std::normal_distribution<double> nd(0.0, 1.0); ref[k] = nd(gen); norm += ref[k] * ref[k]; }
    norm = std::sqrt(norm); for (int k = 0; k < d; k++) ref[k] /= norm;
    uint32_t scratch[624]; double out[8]; wb::random_unit_vector(scratch, seed, v, it, d, out);
    for (int k = 0; k < d; k++) if (out[k] != ref[k]) bad++;
  }
  printf("%d\n", bad); return bad != 0;
}''')
    exe = tmp_path / "mt"
    subprocess.run(["/usr/bin/g++", "-O2", "-I", os.path.join(ROOT, "wembed_b200", "csrc"), str(src), "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)], capture_output=True, text=True).stdout.strip() == "0"


def test_datasets_shapes():
    from wembed_b200.datasets import degree_weights, geometric_graph, heavy_tailed_graph, initial_coordinates
    e, _ = geometric_graph(20000, 10, 1)
    assert 9.0 < 2 * len(e) / 20000 < 11.0 and (e[:, 0] < e[:, 1]).all()
    h, w = heavy_tailed_graph(5000, 20, seed=1)
    deg = np.bincount(h.ravel(), minlength=5000)
    assert 17.0 < deg.mean() < 23.0 and deg.max() > 20 * deg.mean() / 4
    dw = degree_weights(20000, e, 4)
    assert abs(dw.sum() - 20000) < 1e-6
    x = initial_coordinates(100, 3)
    assert (x.astype(np.float32) == x).all() and x.max() < 100 ** (1 / 3) + 1e-6


def test_datagen_helper_matches_numpy_paths():
    """The C++ workload helper (bench / test infrastructure) returns exactly what the numpy code paths return."""
    from wembed_b200 import cabi, datagen
    from wembed_b200.datasets import _pairs_within
    rng = np.random.default_rng(5)
    for n, deg in ((1, 10), (50, 4), (20_000, 10), (5_000, 40)):
        pts = rng.random((n, 2)) * np.sqrt(n)
        r = float(np.sqrt(deg / np.pi))
        a, b = datagen.pairs_within(pts, r), _pairs_within(pts, r)
        assert a.dtype == b.dtype and a.shape == b.shape and (a == b).all()
        rp, col = datagen.csr_canonical(n, a)
        rp2, col2 = cabi.csr_from_edges(n, a)
        assert (rp == rp2).all() and (col == col2).all()
    assert datagen.csr_canonical(3, np.asarray([[1, 0]], np.int32)) is None          # src > dst: not canonical
    assert datagen.csr_canonical(3, np.asarray([[0, 1], [0, 1]], np.int32)) is None  # duplicate


def test_build_tracks_every_kernel_source():
    """A stale library once hid a kernel fix for two GPU runs: every file the translation unit includes must be a build dependency."""
    from wembed_b200 import build
    csrc = os.path.join(ROOT, "wembed_b200", "csrc")
    tracked = {os.path.basename(p) for p in build.DEPS}
    included = set(re.findall(r'#include\s+"([a-z0-9_]+\.cuh)"', "".join(open(os.path.join(csrc, f)).read() for f in os.listdir(csrc))))
    assert included <= tracked and "wb_api.cu" in tracked, included - tracked
