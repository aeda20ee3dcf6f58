"""Edge cases of the device path: empty / tiny graphs, isolated vertices, hubs, duplicate points, round trips."""
import numpy as np
import pytest

import oracle
from helpers import SMALL_GRAPH, lr_exponential, run_to_convergence

pytestmark = pytest.mark.gpu


def test_empty_and_single_vertex(device_lib):
    dev = device_lib.DeviceEmbedder(np.zeros(1, np.int32), np.zeros(0, np.int32))
    st = dev.step(1.0)
    assert st["iteration"] == 1 and st["loss_attract"] == 0.0
    assert dev.coordinates().shape == (0, 4)
    one = device_lib.DeviceEmbedder(np.array([0, 0], np.int32), np.zeros(0, np.int32), embedding_dimension=3)
    one.set_coordinates(np.array([[1.0, 2.0, 3.0]]))
    one.set_weights(np.array([1.0]))
    st = one.step(1.0)                                # WembedEmbedder.cpp:19-21: graphSize() <= 1 -> only the counter moves
    assert st["iteration"] == 1
    np.testing.assert_array_equal(one.coordinates(), [[1.0, 2.0, 3.0]])


def test_two_vertices_against_oracle(device_lib, port_lib):
    for edges, n in (([(0, 1)], 2), ([(0, 1)], 3)):      # n = 3: vertex 2 is isolated
        rp, col = device_lib.csr_from_edges(n, edges)
        x0 = np.array([[0.0, 0.0], [3.0, 4.0], [0.5, 0.1]])[:n]
        w = np.ones(n)
        cpu = oracle.CpuEmbedder("port", edges, n=n, embeddingDimension=2, init_state=False)
        dev = device_lib.DeviceEmbedder(rp, col, embedding_dimension=2, keep_forces=1)
        for e in (cpu, dev):
            e.set_weights(w)
            e.set_coordinates(x0)
        for it in range(1, 6):
            cpu.step()
            st = dev.step(lr_exponential(it))
            assert np.abs(cpu.forces() - dev.forces()).max() <= 1e-5 * max(1e-30, np.abs(cpu.forces()).max())
            assert np.abs(cpu.coordinates() - dev.coordinates()).max() <= 1e-5 * max(1.0, np.abs(cpu.coordinates()).max())
            np.testing.assert_allclose(st["loss_attract"], cpu.stats()["loss_attract"], rtol=1e-5, atol=1e-7)
            cpu.set_coordinates(dev.coordinates())     # one trajectory: fp32 rounding must not feed back through the dynamics


def test_star_hub_and_isolated(device_lib, port_lib):
    """One hub adjacent to everything (row length n-1), the rest leaves, plus isolated vertices."""
    n, d = 5000, 4
    edges = [(0, v) for v in range(1, n - 50)]
    rng = np.random.default_rng(0)
    x0 = (rng.random((n, d)) * 4).astype(np.float32).astype(np.float64)
    from wembed_b200.datasets import degree_weights
    w = degree_weights(n, np.asarray(edges), d)
    rp, col = device_lib.csr_from_edges(n, edges)
    cpu = oracle.CpuEmbedder("port", edges, n=n, embeddingDimension=d, init_state=False)
    dev = device_lib.DeviceEmbedder(rp, col, embedding_dimension=d, keep_forces=1)
    for e in (cpu, dev):
        e.set_weights(w)
        e.set_coordinates(x0)
    for it in range(1, 4):
        cpu.step()
        st = dev.step(lr_exponential(it))
        fr, fd = cpu.forces(), dev.forces()
        # the hub sums 4949 terms in fp32: tolerance relative to the sum of magnitudes, 1e-5 elsewhere
        assert np.abs(fr[1:] - fd[1:]).max() <= 1e-5 * np.abs(fr).max()
        assert np.abs(fr[0] - fd[0]).max() <= 1e-5 * np.abs(fr).max()
        assert abs(st["num_repulsion_pairs"] - cpu.stats()["num_rep_pairs"]) <= 8
        cpu.set_coordinates(dev.coordinates())


def test_duplicate_points_use_tie_break(device_lib, port_lib):
    """Pairs of exactly coincident vertices (neighbours and non-neighbours) inside an otherwise generic layout."""
    n, d = 400, 3
    rng = np.random.default_rng(2)
    edges = [(i, i + 1) for i in range(n - 1)] + [(i, i + 5) for i in range(n - 5)]
    x0 = (rng.random((n, d)) * 5).astype(np.float32).astype(np.float64)
    x0[10] = x0[11]          # coincident neighbours  -> attraction tie-break
    x0[100] = x0[300]        # coincident non-neighbours -> repulsion tie-break
    x0[200] = x0[201] = x0[350]
    w = np.ones(n)
    rp, col = device_lib.csr_from_edges(n, edges)
    cpu = oracle.CpuEmbedder("port", edges, n=n, seed=1234, embeddingDimension=d, init_state=False)
    dev = device_lib.DeviceEmbedder(rp, col, embedding_dimension=d, keep_forces=1, seed=1234)
    for e in (cpu, dev):
        e.set_weights(w)
        e.set_coordinates(x0)
    cpu.step()
    dev.step(lr_exponential(1))
    fr, fd = cpu.forces(), dev.forces()
    assert np.abs(fr - fd).max() <= 1e-5 * np.abs(fr).max()
    assert np.abs(fr[[10, 11, 100, 300, 200, 201, 350]]).max() > 0


def test_coordinate_and_weight_round_trip(device_lib):
    n, d = 1000, 5
    rng = np.random.default_rng(4)
    edges = [(i, (i + 1) % n) for i in range(n)]
    rp, col = device_lib.csr_from_edges(n, edges)
    dev = device_lib.DeviceEmbedder(rp, col, embedding_dimension=d)
    x = rng.normal(size=(n, d)) * 100
    w = rng.random(n) + 0.1
    dev.set_coordinates(x)
    dev.set_weights(w)
    np.testing.assert_array_equal(dev.coordinates(), x.astype(np.float32).astype(np.float64))
    np.testing.assert_array_equal(dev.weights(), w)
    with pytest.raises(device_lib.WbError):
        dev.set_weights(np.zeros(n))
    with pytest.raises(device_lib.WbError):
        dev.forces()           # keep_forces was not requested


def test_small_graph_reaches_zero_loss(device_lib):
    """BASELINE.json configs[0]: assets/small_graph.edg, default options -> converges to loss exactly 0 like the reference."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "small_graph_seed1.npz"))
    rp, col = device_lib.csr_from_edges(5, SMALL_GRAPH)
    np.testing.assert_array_equal(rp, g["csr_row"])
    np.testing.assert_array_equal(col, g["csr_col"])
    dev = device_lib.DeviceEmbedder(rp, col, embedding_dimension=4, seed=1)
    dev.set_weights(g["w"])
    dev.set_coordinates(g["x0"])
    iters, st = run_to_convergence(dev, {})
    assert st["loss_attract"] + st["loss_repel"] == 0.0
    assert 0.5 * int(g["iterations"]) <= iters <= 2.0 * int(g["iterations"]), (iters, int(g["iterations"]))


@pytest.mark.parametrize("d,lo,hi,rep", [(2, 1e-3, 1e3, 1.0), (1, 1e-2, 1e2, 1.0), (3, 1e-4, 1.0, 250.0), (8, 1.0, 1e6, 1e-3)])
def test_wide_weight_ranges_fixed_point_rows(device_lib, port_lib, d, lo, hi, rep):
    """The repulsion rows are 64-bit fixed point with a scale derived from the weights and repulsionScale
    (wb_set_weights): forces and losses must keep the 1e-5 parity for weight ranges of six decades."""
    n = 3000
    rng = np.random.default_rng(11)
    edges = np.asarray([(i, (i + 1) % n) for i in range(n)] + [(i, (i + 17) % n) for i in range(n)], np.int32)
    w = np.exp(rng.uniform(np.log(lo), np.log(hi), n))
    side = 3.0 * float(np.median(w)) ** (2.0 / d) * n ** (1.0 / d) / 4.0     # a few partners per vertex at the median radius
    x0 = (rng.random((n, d)) * max(side, 1.0)).astype(np.float32).astype(np.float64)
    rp, col = device_lib.csr_from_edges(n, edges)
    cpu = oracle.CpuEmbedder("port", edges, n=n, embeddingDimension=d, init_state=False, repulsionScale=rep)
    dev = device_lib.DeviceEmbedder(rp, col, embedding_dimension=d, keep_forces=1, repulsion_scale=rep)
    for e in (cpu, dev):
        e.set_weights(w)
        e.set_coordinates(x0)
    cpu.step()
    st = dev.step(lr_exponential(1))
    cs = cpu.stats()
    fr, fd = cpu.forces(), dev.forces()
    from helpers import near_threshold_vertices
    ok = ~near_threshold_vertices(x0, w, rp, col)
    assert ok.mean() > 0.95
    assert cs["num_rep_pairs"] > 0
    assert np.abs(fr[ok] - fd[ok]).max() <= 1e-5 * np.abs(fr).max()
    assert abs(st["num_repulsion_pairs"] - cs["num_rep_pairs"]) <= 2 * (~ok).sum()
    if ok.all():
        np.testing.assert_allclose(st["loss_repel"], cs["loss_repel"], rtol=1e-5)
        np.testing.assert_allclose(st["loss_attract"], cs["loss_attract"], rtol=1e-5)


@pytest.mark.parametrize("d", [1, 2, 4, 8, 11, 16])
def test_half_precision_boxes_find_the_same_pairs(device_lib, monkeypatch, d):
    """The box rounds of the repulsion walk may test a half-precision copy of the boxes (k_repulse_pairs<V, true>); it only
    prunes, the pair test stays exact fp32, and the rows are integer sums: forced on and forced off must agree BIT FOR BIT in
    forces, coordinates and pair counts - on a blob, a layout 300 radii wide, outliers beyond the half range (1e5, where
    coordinates round to inf) and weights spanning four decades (radii beyond what half precision can square)."""
    n = 6000
    edges, _ = geometric_graph_small(n)
    rp, col = device_lib.csr_from_edges(n, edges)
    rng = np.random.default_rng(d)
    blob = rng.normal(0.0, {1: 1.5, 2: 1.5, 4: 1.2, 8: 0.5, 11: 0.35, 16: 0.25}[d], (n, d))   # dense enough for repulsive pairs
    wide = blob * 200.0
    outliers = blob.copy()
    outliers[rng.choice(n, 40, replace=False)] *= 1.0e5
    outliers[rng.choice(n, 5, replace=False), 0] = 7.0e4
    shifted = blob + 3000.0                                # far from the origin: the frame centre has to absorb it
    from wembed_b200.datasets import degree_weights
    w_deg = degree_weights(n, edges, d)
    w_wide = np.exp(rng.uniform(np.log(1e-2), np.log(1e2), n))
    w_wide *= n / w_wide.sum()
    for name, x0, w in (("blob", blob, w_deg), ("wide", wide, w_deg), ("outliers", outliers, w_deg), ("shifted", shifted, w_deg),
                        ("weights", blob, w_wide)):
        x0 = x0.astype(np.float32).astype(np.float64)
        out = {}
        for mode in ("0", "1"):
            monkeypatch.setenv("WB_HALF_BOXES", mode)
            dev = device_lib.DeviceEmbedder(rp, col, embedding_dimension=d, keep_forces=1, seed=7)
            dev.set_weights(w)
            dev.set_coordinates(x0)
            pairs = []
            for it in range(1, 4):
                pairs.append(dev.step(lr_exponential(it))["num_repulsion_pairs"])
            out[mode] = (pairs, dev.forces().copy(), dev.coordinates().copy())
        assert out["0"][0] == out["1"][0], (name, out["0"][0], out["1"][0])
        np.testing.assert_array_equal(out["0"][1], out["1"][1], err_msg=name)
        np.testing.assert_array_equal(out["0"][2], out["1"][2], err_msg=name)
        assert name not in ("blob", "shifted", "weights") or out["0"][0][0] > 0, name   # the comparison is not vacuous


def geometric_graph_small(n):
    from wembed_b200.datasets import geometric_graph
    return geometric_graph(n, 10, 11)


def test_pair_list_policy_never_changes_results(device_lib):
    """The repulsion pair list may be kept for several steps (skin > 0) or rebuilt every step (skin_max = 0, what
    WembedEmbedder::updateIndex does): a listed pair beyond the exact threshold contributes exactly zero and the sums are integers,
    so the whole trajectory is bit-identical - and late in the run most steps must have reused their list."""
    from helpers import make_problem
    n, d, steps = 3000, 2, 420
    edges, w, x0 = make_problem(n, d)
    rp, col = device_lib.csr_from_edges(n, edges)
    out = []
    for policy in ((0.0, 4.0), (1.0, 4.0), (0.3, 8.0)):
        dev = device_lib.DeviceEmbedder(rp, col, embedding_dimension=d, seed=1234)
        dev.set_list_policy(*policy)
        dev.set_weights(w)
        dev.set_coordinates(x0)
        # (a faster cooling than the default 0.995, so the layout settles - and lists start to live - within the test's steps)
        stats = [dev.step(lr_exponential(it, cooling=0.97)) for it in range(1, steps + 1)]
        out.append((dev.coordinates(), [(s["loss_attract"], s["loss_repel"], s["num_repulsion_pairs"], s["sum_displacement"]) for s in stats],
                    sum(1 for s in stats if s["list_rebuilt"] == 0), max(s["list_skin"] for s in stats)))
        dev.close()
    assert out[0][2] == 0 and out[0][3] == 0.0                      # skin_max = 0: every step searched
    for x, s, reused, skin in out[1:]:
        assert np.array_equal(x, out[0][0])
        assert s == out[0][1]
        assert reused > steps // 4 and skin > 0.0


def test_pair_list_grows_on_overflow(device_lib, monkeypatch):
    """A pair buffer that is too small stops the step on the device (nothing is applied), the host grows it and replays the queued
    steps: same results as with a large buffer, for blocking and for queued steps."""
    from helpers import make_problem
    n, d = 4000, 3
    edges, w, x0 = make_problem(n, d)
    rp, col = device_lib.csr_from_edges(n, edges)
    out = {}
    for cap, queued in (("0", False), ("16", False), ("16", True)):
        if cap != "0":
            monkeypatch.setenv("WB_PAIR_CAP", cap)
        else:
            monkeypatch.delenv("WB_PAIR_CAP", raising=False)
        dev = device_lib.DeviceEmbedder(rp, col, embedding_dimension=d, seed=1234)
        dev.set_weights(w)
        dev.set_coordinates(x0)
        if queued:
            for it in range(1, 13):
                dev.step_async(lr_exponential(it))
            stats = [dev.step_collect() for _ in range(12)]
        else:
            stats = [dev.step(lr_exponential(it)) for it in range(1, 13)]
        out[(cap, queued)] = (dev.coordinates(), [(s["loss_attract"], s["loss_repel"], s["num_repulsion_pairs"]) for s in stats])
        dev.close()
    base = out[("0", False)]
    assert base[1][0][2] > 16                                       # the first step alone lists more pairs than the tiny buffer holds
    for key in (("16", False), ("16", True)):
        assert np.array_equal(out[key][0], base[0])
        assert out[key][1] == base[1]


def test_graph_replay_equals_direct_launches(device_lib, monkeypatch):
    """One GPU, no phase timing: a step is one CUDA graph launch (k_step_begin -> IF(rebuild){index, search, pair list} -> forces .. tail).
    Same kernels, same arguments: the trajectory must be bit-identical to direct launches, through rebuild and reuse steps, blocking and
    queued, and across a re-capture (wb_set_weights voids the captured arguments)."""
    from helpers import make_problem
    n, d, steps = 5000, 4, 260
    edges, w, x0 = make_problem(n, d)
    rp, col = device_lib.csr_from_edges(n, edges)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("WB_GRAPH", mode)
        dev = device_lib.DeviceEmbedder(rp, col, embedding_dimension=d, seed=1234)
        dev.set_weights(w)
        dev.set_coordinates(x0)
        stats = [dev.step(lr_exponential(it, cooling=0.97)) for it in range(1, steps // 2 + 1)]
        dev.set_weights(w)                                   # forces a new capture
        inflight = 0
        for it in range(steps // 2 + 1, steps + 1):
            dev.step_async(lr_exponential(it, cooling=0.97))
            inflight += 1
            if inflight == 16:
                stats.extend(dev.step_collect() for _ in range(16))
                inflight = 0
        stats.extend(dev.step_collect() for _ in range(inflight))
        out[mode] = (dev.coordinates(), [(s["loss_attract"], s["loss_repel"], s["num_repulsion_pairs"], s["list_rebuilt"]) for s in stats], dev.exec_mode())
        dev.close()
    assert out["1"][2][0] == "graph", out["1"][2]
    assert out["0"][2][0] == "direct"
    assert np.array_equal(out["1"][0], out["0"][0])
    assert out["1"][1] == out["0"][1]
    assert any(s[3] == 0 for s in out["1"][1]) and any(s[3] == 1 for s in out["1"][1])     # both kinds of step were replayed
