"""Multilevel driver: host-side coarsening against the reference's parent pointers (golden fixture
generated from the reference's LabelPropagation), and the layered embedding through the public API on the GPU."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "hierarchy.npz"))


@pytest.fixture(scope="module")
def host_lib():
    from wembed_b200 import cabi, host
    host.build()
    cabi.lib()                                   # libwembed_host.so depends on libwembed_b200.so
    C.CDLL(cabi.LIB_PATH, mode=C.RTLD_GLOBAL)
    return C.CDLL(host.HOST_LIB)


@pytest.mark.parametrize("name", ["ring64", "geo3000", "heavy4000"])
def test_coarsening_matches_reference(host_lib, name):
    """LabelPropagation::coarsenAllLayers (LabelPropagation.cpp:13-56) incl. the aggressive fallback and coarsenGraph:
    identical layer sizes and identical parent pointers (integer work: bit-exact)."""
    e = GOLD[f"{name}_edges"]
    src, dst = np.ascontiguousarray(e[:, 0], dtype=np.int32), np.ascontiguousarray(e[:, 1], dtype=np.int32)
    sizes, parents = np.zeros(64, np.int32), np.full(4 * len(e) + 64, -7, np.int32)
    ip = C.POINTER(C.c_int32)
    host_lib.wbh_coarsen.restype = C.c_int
    nl = host_lib.wbh_coarsen(C.c_longlong(len(src)), src.ctypes.data_as(ip), dst.ctypes.data_as(ip), sizes.ctypes.data_as(ip), 64,
                              parents.ctypes.data_as(ip), C.c_longlong(len(parents)))
    np.testing.assert_array_equal(sizes[:nl], GOLD[f"{name}_sizes"])
    np.testing.assert_array_equal(parents[: sizes[:nl].sum()], GOLD[f"{name}_parents"])
    assert sizes[nl - 1] == 1 and parents[sizes[:nl].sum() - 1] == -1


@pytest.mark.gpu
def test_layered_embedding_through_public_api():
    """Options::layeredEmbedding = true: coarsest layer first, expansion on convergence, same final quality as the reference's
    LayeredEmbedder run on this graph (golden: loss 0, constructDeg = MAP = 1)."""
    from helpers import reconstruction_metrics
    from wembed_b200 import cabi, host
    wembed = host.load()
    e = GOLD["geo3000_edges"]
    n = int(e.max()) + 1
    wembed.setSeed(7)
    o = wembed.Options()
    o.layeredEmbedding = True
    graph = wembed.graphFromEdges([wembed.Edge(int(a), int(b)) for a, b in e])
    emb = wembed.createEmbedder(graph, o)
    assert emb.getNumVertices() == 1 and emb.getCurrentGraph().getNumVertices() == 1     # starts at the single-vertex layer
    sizes = [emb.getNumVertices()]
    steps = 0
    while not emb.isFinished():
        emb.calculateStep()
        steps += 1
        if emb.getNumVertices() != sizes[-1]:
            sizes.append(emb.getNumVertices())
        assert steps < 40000
    assert sizes == list(GOLD["geo3000_sizes"][::-1])
    ref_iters, ref_loss, ref_cd, ref_map = GOLD["geo3000_layered"]
    assert 0.5 * ref_iters <= steps <= 2.0 * ref_iters, (steps, ref_iters)
    x, w = np.asarray(emb.getCoordinates()), np.asarray(emb.getWeights())
    rp, col = cabi.csr_from_edges(n, e)
    cd, mp = reconstruction_metrics(x, w, rp, col, np.arange(0, n, 6))
    assert abs(cd - ref_cd) <= 0.01 * ref_cd and abs(mp - ref_map) <= 0.01 * ref_map, (cd, mp)
    assert emb.getLoss().total <= 1e-3 * n
    names = [t.display_name for t in emb.getTimings()]
    assert "Expanding Positions" in names
    emb.setCoordinates(x.tolist())            # warns, no effect (LayeredEmbedder.cpp:26-30)
