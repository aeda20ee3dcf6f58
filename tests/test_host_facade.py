"""The host-side mirror of the reference interface (include/wembed.h + the pybind11 module `wembed`)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_HEADER = "/root/reference/include/wembed.h"
REF_BINDINGS = "/root/reference/python/bindings.cpp"


@pytest.fixture(scope="module")
def wembed():
    from wembed_b200 import host
    return host.load()


def test_module_surface_matches_reference_bindings(wembed):
    """Every name the reference's python/bindings.cpp:11-133 exports exists here."""
    expected = {"SpatialIndex", "IndexSNN", "IndexSprk", "OptimizerType", "OptimizerSimple", "OptimizerAdam", "LRSchedule",
                "LRExponentialCooling", "LRLossAdaptive", "StopCriterion", "StopDisplacement", "StopLoss", "Edge", "TimingResult",
                "Loss", "Options", "Graph", "Embedder", "createEmbedder", "graphFromEdges", "graphFromEdgeListFile",
                "readCoordinatesFromFile", "timingsToString", "setSeed", "__version__"}
    assert expected <= set(dir(wembed))
    if os.path.exists(REF_BINDINGS):
        src = open(REF_BINDINGS).read()
        for name in re.findall(r'm\.def\("(\w+)"', src) + re.findall(r'py::class_<[^>]+>\(m, "(\w+)"\)', src) + re.findall(r'\.value\("(\w+)"', src):
            assert hasattr(wembed, name), name
        for cls, pat in (("Embedder", r'py::class_<wembed::Embedder>(.*?);'), ("Graph", r'py::class_<wembed::Graph>(.*?);')):
            body = re.search(pat, src, re.S).group(1)
            for meth in re.findall(r'\.def\("(\w+)"', body):
                assert hasattr(getattr(wembed, cls), meth), (cls, meth)


def test_options_fields_and_defaults_match_reference_header(wembed):
    o = wembed.Options()
    defaults = dict(embeddingDimension=4, useUnitWeights=False, dimensionHint=-1.0, layeredEmbedding=False, attractionScale=1.0,
                    repulsionScale=1.0, centreScale=0.0, edgeLength=1.0, expansionStretch=1.0, maxIterations=10000,
                    simpleOptMaxDisplacement=1.0, learningRate=10.0, warmupSteps=20, lrCoolingFactor=0.995, lrDecayFactor=0.5,
                    lrDecayThreshold=1e-2, lrAdaptPatience=20, lrGrowthFactor=1.0, lrGrowthThreshold=1e-1, stopDisplacementTol=3e-4,
                    stopDisplacementPatience=5, lossSmoothingFactor=0.3, lossRateWindow=30, stopLossTol=1e-3, stopLossPatience=50)
    for k, v in defaults.items():
        assert getattr(o, k) == v, k
    assert (o.indexType, o.optimizerType, o.lrSchedule, o.stopCriterion) == (wembed.IndexSprk, wembed.OptimizerAdam, wembed.LRExponentialCooling, wembed.StopLoss)
    assert (int(wembed.IndexSNN), int(wembed.IndexSprk), int(wembed.OptimizerSimple), int(wembed.OptimizerAdam)) == (1, 2, 0, 1)
    if os.path.exists(REF_HEADER):
        body = re.search(r"struct Options \{(.*?)\n\};", open(REF_HEADER).read(), re.S).group(1)
        fields = re.findall(r"^\s*[\w:]+\s+(\w+)\s*=\s*([^;]+);", body, re.M)
        assert len(fields) == 29
        for name, default in fields:
            assert hasattr(o, name), name
            if re.fullmatch(r"-?[\d.e+-]+", default.strip()):
                assert float(getattr(o, name)) == float(default), name


def test_graph_kats(wembed):
    """tests/TestGraph.cpp:22-29,61-139 through the public API."""
    E = wembed.Edge
    g = wembed.graphFromEdges([E(0, 1), E(0, 2), E(1, 2), E(2, 3), E(3, 0), E(0, 1), E(1, 0), E(2, 2)])
    assert (g.getNumVertices(), g.getNumEdges()) == (4, 5)
    assert [g.getNeighbors(v) for v in range(4)] == [[1, 2, 3], [0, 2], [0, 1, 3], [0, 2]]
    assert [g.getNumNeighbors(v) for v in range(4)] == [3, 2, 3, 2]
    assert g.getEdges(2) == [5, 6, 7] and [g.getEdgeTarget(e) for e in g.getEdges(2)] == [0, 1, 3]
    assert g.areNeighbors(0, 3) and g.areNeighbors(3, 0) and not g.areNeighbors(1, 3) and not g.areNeighbors(2, 2)
    assert [(e.src, e.dst) for e in g.getEdgeList()] == [(0, 1), (0, 2), (0, 3), (1, 2), (2, 3)]
    assert repr(g).startswith("Graph AdjList:\n0: 1 2 3 \n")


def test_edge_list_and_coordinate_files(wembed, tmp_path):
    p = tmp_path / "g.edg"
    p.write_text("# This line will be ignored\n# n=5, m=7\n0 1\n1 2\n2 3\n3 4\n1 3\n2 4\n")   # assets/small_graph.edg
    g = wembed.graphFromEdgeListFile(str(p))
    assert (g.getNumVertices(), g.getNumEdges()) == (5, 6)
    assert g.getNeighbors(1) == [0, 2, 3]
    q = tmp_path / "g.csv"
    q.write_text("% comment\n0;1\n1;2\n")
    assert wembed.graphFromEdgeListFile(str(q), "%", ";").getNumEdges() == 2
    c = tmp_path / "coords.csv"
    c.write_text("% header\n1,0.5,1.5,2.0\n0,-1.0,2.25,3.0\n")
    assert wembed.readCoordinatesFromFile(str(c)) == [[-1.0, 2.25, 3.0], [0.5, 1.5, 2.0]]
    with pytest.raises(RuntimeError):
        wembed.graphFromEdgeListFile(str(tmp_path / "missing.edg"))


def test_timings_to_string(wembed):
    assert wembed.timingsToString([]) == ""


@pytest.mark.gpu
def test_embed_small_graph_through_public_api(wembed, tmp_path):
    """BASELINE.json configs[0] through the drop-in API: createEmbedder -> calculateEmbedding -> getCoordinates."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "small_graph_seed1.npz"))
    wembed.setSeed(1)
    graph = wembed.graphFromEdges([wembed.Edge(int(a), int(b)) for a, b in g["edges"]])
    o = wembed.Options()
    o.indexType = wembed.IndexSNN
    emb = wembed.createEmbedder(graph, o)
    # same initial layout (Rand::randomCoordinates from mt19937(1)) and degree weights as the reference
    np.testing.assert_allclose(np.asarray(emb.getCoordinates()), g["x0"].astype(np.float32), rtol=1e-7)
    np.testing.assert_allclose(emb.getWeights(), g["w"], rtol=1e-15)
    assert not emb.isFinished() and emb.getCurrentLearningRate() == 10.0
    emb.calculateStep()
    assert emb.getCurrentLearningRate() == pytest.approx(10.0 * 0.995 / 20.0)
    emb.calculateEmbedding()
    assert emb.isFinished()
    loss = emb.getLoss()
    assert loss.total == 0.0 == loss.attractive + loss.repulsive
    x = np.asarray(emb.getCoordinates())
    assert x.shape == (5, 4) and emb.getNumVertices() == 5 and emb.getEmbeddingDimension() == 4
    assert np.abs(x.mean(axis=0)).max() < 1e-5
    # a valid embedding: neighbours within, non-neighbours beyond the weighted threshold
    w = np.asarray(emb.getWeights())
    for a in range(5):
        for b in range(a + 1, 5):
            dw = np.linalg.norm(x[a] - x[b]) / (w[a] * w[b]) ** 0.25
            assert (dw <= 1.0 + 1e-5) if graph.areNeighbors(a, b) else (dw >= 1.0 - 1e-5)
    names = [t.display_name for t in emb.getTimings()]
    assert "Embedding" in names and "Compute Repelling Forces" in names and "Construct spacial index" in names
    assert "Embedding" in wembed.timingsToString(emb.getTimings())
    out = tmp_path / "x.csv"
    emb.writeCoordinates(str(out))
    back = np.asarray(wembed.readCoordinatesFromFile(str(out)))
    np.testing.assert_allclose(back[:, :4], x, rtol=1e-15)
    np.testing.assert_allclose(back[:, 4], w, rtol=1e-15)
    emb.setCoordinates(back[:, :4].tolist())
    emb.setWeights(w.tolist())
    assert emb.getCurrentGraph().getNumEdges() == 6


@pytest.mark.gpu
def test_ring64_stop_criteria_through_public_api(wembed):
    """tests/TestDeterminism.cpp:112-165 via the public API: both stop criteria fire before maxIterations and the
    run is reproducible for a fixed seed."""
    from helpers import ring_graph
    edges = [wembed.Edge(int(a), int(b)) for a, b in ring_graph(64)]
    results = []
    for _ in range(2):
        wembed.setSeed(1234)
        o = wembed.Options()
        o.embeddingDimension = 2
        o.stopCriterion = wembed.StopDisplacement
        o.stopDisplacementTol = 1e-3
        o.stopDisplacementPatience = 5
        o.maxIterations = 5000
        emb = wembed.createEmbedder(wembed.graphFromEdges(edges), o)
        steps = 0
        while not emb.isFinished():
            emb.calculateStep()
            steps += 1
        results.append((steps, np.asarray(emb.getCoordinates())))
    assert 5 < results[0][0] < 5000
    assert results[0][0] == results[1][0] and np.array_equal(results[0][1], results[1][1])
    wembed.setSeed(1234)
    o = wembed.Options()
    o.embeddingDimension = 2
    o.lrSchedule = wembed.LRLossAdaptive
    o.lossRateWindow = 10
    o.lrAdaptPatience = 5
    o.maxIterations = 1000
    emb = wembed.createEmbedder(wembed.graphFromEdges(edges), o)
    rates = []
    for _ in range(300):
        emb.calculateStep()
        rates.append(emb.getCurrentLearningRate())
    assert rates[0] == pytest.approx(10.0 / 20.0) and max(rates) == 10.0      # warm-up ramp, then the initial rate
    assert min(rates[30:]) <= 5.0                                              # a plateau decay fired (TestDeterminism.cpp:131-147)
    assert all(b <= a for a, b in zip(rates[20:], rates[21:]))                 # growth is off: the rate never increases


@pytest.mark.parametrize("case", ["loss_stop", "disp_stop", "adaptive", "default"])
def test_host_scalar_logic_matches_reference(case):
    """LRScheduler (both schedules + warm-up), ConvergenceMonitor, DisplacementMonitor and isFinished of the C++ facade, replayed
    over the reference's own loss / displacement sequences (tests/TestDeterminism.cpp option sets on ring-64; golden from the
    reference build): learning rates, loss rates and the stop iteration must be bit-identical - it is pure double arithmetic."""
    import ctypes as C
    from wembed_b200 import cabi, host
    host.build()
    cabi.lib()
    C.CDLL(cabi.LIB_PATH, mode=C.RTLD_GLOBAL)
    lib = C.CDLL(host.HOST_LIB)
    g = np.load(os.path.join(ROOT, "tests", "golden", "host_logic.npz"))
    trace, opts = g[f"{case}_trace"], np.ascontiguousarray(g[f"{case}_opts"])
    steps = len(trace)
    loss, disp = np.ascontiguousarray(trace[:, 0]), np.ascontiguousarray(trace[:, 1])
    lr, rate = np.zeros(steps), np.zeros(steps)
    dp = C.POINTER(C.c_double)
    lib.wbh_host_logic_trace.restype = C.c_int
    stop = lib.wbh_host_logic_trace(opts.ctypes.data_as(dp), steps, loss.ctypes.data_as(dp), disp.ctypes.data_as(dp), lr.ctypes.data_as(dp),
                                    rate.ctypes.data_as(dp))
    np.testing.assert_array_equal(lr, trace[:, 2])
    np.testing.assert_array_equal(rate, trace[:, 3])          # inf during the window warm-up, then the windowed relative decrease
    assert stop == steps                                      # the reference stopped exactly after its last recorded step
    if case == "adaptive":
        assert lr.min() < 10.0 and (np.diff(lr[20:]) <= 0).all()


def test_edge_list_file_ingestion_matches_the_reference_at_scale(ref_lib, tmp_path):
    """wembed::graphFromEdgeListFile (sort-based CSR build, wembed_b200/host/graph.cpp) against the reference's own
    GraphIO::readEdgeList + Graph(std::map<int, std::set<int>>) (GraphIO.cpp:10-51, Graph.cpp:87-150) on a generated file with one
    million edge lines: comments, both orientations, repeated lines and one self loop.  The CSR must be identical; both are timed."""
    import time
    import oracle
    from wembed_b200 import host
    from wembed_b200.datasets import geometric_graph
    wembed = host.load()
    n = 200_000
    edges, _ = geometric_graph(n, 10, seed=11)
    rng = np.random.default_rng(5)
    lines = edges[rng.permutation(len(edges))].astype(np.int64)
    flip = rng.random(len(lines)) < 0.5
    lines[flip] = lines[flip][:, ::-1]                                   # either orientation
    lines = np.concatenate([lines, lines[:20_000], [[7, 7]]])            # repeated lines, one self loop
    path = tmp_path / "graph.edg"
    with open(path, "w") as f:
        f.write("# generated edge list\\n# n m\\n")
        np.savetxt(f, lines, fmt="%d", delimiter=" ")
    assert len(lines) > 1_000_000
    t0 = time.perf_counter()
    g = wembed.graphFromEdgeListFile(str(path))
    t_ours = time.perf_counter() - t0
    rp, col = g.csr()
    t0 = time.perf_counter()
    rp_ref, col_ref = oracle.ref_read_edge_list(path)
    t_ref = time.perf_counter() - t0
    np.testing.assert_array_equal(rp, rp_ref)
    np.testing.assert_array_equal(col, col_ref)
    assert g.getNumVertices() == len(rp_ref) - 1 and g.getNumEdges() == len(col_ref) // 2 == len(edges)
    # the array constructor takes the same route
    g2 = wembed.graphFromEdgeArray(lines.astype(np.int32))
    np.testing.assert_array_equal(g2.csr()[1], col_ref)
    print(f"edge-list ingestion, {len(lines)} lines: this build {t_ours:.2f} s, reference {t_ref:.2f} s")
