"""north_star: "Converged-embedding quality (evaluationLib reconstruction metrics) must match to within 1%"."""
import numpy as np
import pytest

import oracle
from helpers import edge_detection_metrics, lr_exponential, make_problem, reconstruction_metrics, run_to_convergence


def test_reconstruction_metric_kat():
    """tests/TestMetrics.cpp:29-62: 3-vertex path, good and bad Euclidean coordinates (unit weights)."""
    rp, col = np.array([0, 1, 3, 4]), np.array([1, 0, 2, 1])
    w = np.ones(3)
    good = np.array([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0]])
    assert reconstruction_metrics(good, w, rp, col) == (1.0, 1.0)
    bad = np.array([[0.0, 0.0], [3.0, 0.0], [1.0, 0.0]])
    cd, mp = reconstruction_metrics(bad, w, rp, col)
    assert cd == pytest.approx(1.0 / 3.0) and mp == pytest.approx((0.5 + 0.5 + 1.0) / 3.0)


@pytest.mark.gpu
def test_quality_matches_oracle_within_one_percent(device_lib, port_lib):
    """Fixed iteration budget (the loss stop fires late on perfectly embeddable graphs) and each side's own
    stop: constructDeg and MAP of the device embedding vs the CPU oracle's, same graph, same initial layout."""
    n, d = 3000, 4
    edges, w, x0 = make_problem(n, d)
    rp, col = device_lib.csr_from_edges(n, edges)
    nodes = np.random.default_rng(0).choice(n, 500, replace=False)      # cli_evaluator samples <= 1000 nodes
    budget = 400
    cpu = oracle.CpuEmbedder("port", edges, n=n, embeddingDimension=d, init_state=False)
    dev = device_lib.DeviceEmbedder(rp, col, embedding_dimension=d, seed=1234)
    for e in (cpu, dev):
        e.set_weights(w)
        e.set_coordinates(x0)
    for it in range(1, budget + 1):
        cpu.step()
        st = dev.step(lr_exponential(it))
    q_cpu = reconstruction_metrics(cpu.coordinates(), w, rp, col, nodes)
    q_dev = reconstruction_metrics(dev.coordinates(), w, rp, col, nodes)
    assert q_cpu[0] > 0.5 and q_cpu[1] > 0.5
    for a, b in zip(q_dev, q_cpu):
        assert abs(a - b) <= 0.01 * b, (q_dev, q_cpu)
    # losses follow the same curve (chaotic in detail, equal in aggregate)
    total_cpu = cpu.stats()["loss_attract"] + cpu.stats()["loss_repel"]
    assert st["loss_attract"] + st["loss_repel"] == pytest.approx(total_cpu, rel=0.15)
    # to each side's own stop (default loss criterion)
    it_dev, _ = run_to_convergence(dev, {"maxIterations": 4000})
    q_final = reconstruction_metrics(dev.coordinates(), w, rp, col, nodes)
    assert q_final[0] >= q_dev[0] - 0.01 and q_final[1] >= q_dev[1] - 0.01


@pytest.mark.gpu
def test_device_reconstruction_metric(device_lib):
    """wb_reconstruction against the reference's known answers (tests/TestMetrics.cpp:29-62) and against the numpy
    restatement of NodeSampler on a geometric graph with hubs and isolated vertices."""
    rp, col = np.array([0, 1, 3, 4], np.int32), np.array([1, 0, 2, 1], np.int32)
    dev = device_lib.DeviceEmbedder(rp, col, embedding_dimension=2)
    dev.set_weights(np.ones(3))
    dev.set_coordinates(np.array([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0]]))
    assert dev.reconstruction([0, 1, 2]) == (1.0, 1.0)
    dev.set_coordinates(np.array([[0.0, 0.0], [3.0, 0.0], [1.0, 0.0]]))
    cd, mp = dev.reconstruction([0, 1, 2])
    assert cd == pytest.approx(1.0 / 3.0) and mp == pytest.approx((0.5 + 0.5 + 1.0) / 3.0)

    n, d = 4000, 5
    edges, w, x0 = make_problem(n, d)
    edges = np.concatenate([edges[(edges != 17).all(axis=1)], [(3, v) for v in range(100, 2300)]])   # 17 isolated, 3 a hub
    from wembed_b200.datasets import degree_weights
    w = degree_weights(n, edges, d)
    rp, col = device_lib.csr_from_edges(n, edges)
    dev = device_lib.DeviceEmbedder(rp, col, embedding_dimension=d, seed=1)
    dev.set_weights(w)
    dev.set_coordinates(x0)
    for it in range(1, 60):
        dev.step(lr_exponential(it))
    nodes = np.concatenate([[3, 17], np.random.default_rng(1).choice(n, 300, replace=False)])
    got = dev.reconstruction(nodes)
    exp = reconstruction_metrics(dev.coordinates(), w, rp, col, nodes)
    assert got == pytest.approx(exp, rel=1e-9), (got, exp)
    assert 0.0 < got[0] < 1.0


def _metric_kat_inputs():
    """tests/TestMetrics.cpp:64-93: path 0-1-2, edgeSampleScale 1 (=> the one non-edge is sampled with probability 1)."""
    rp, col = np.array([0, 1, 3, 4], np.int32), np.array([1, 0, 2, 1], np.int32)
    from wembed_b200.metrics import sample_edge_pairs
    v, u, f = sample_edge_pairs(rp, col, scale=1.0, seed=0)
    assert sorted(zip(v.tolist(), u.tolist(), f.tolist())) == [(0, 1, 1), (0, 2, 0), (1, 2, 1)]
    good = np.array([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0]])
    bad = np.array([[0.0, 0.0], [3.0, 0.0], [1.0, 0.0]])
    return rp, col, v, u, f, good, bad


def test_edge_detection_metric_kat():
    """The numpy restatement of EdgeDetection against the reference's known answers, and the sampler's invariants."""
    rp, col, v, u, f, good, bad = _metric_kat_inputs()
    w = np.ones(3)
    assert edge_detection_metrics(good, w, v, u, f, 3, 2) == pytest.approx((1.0, 1.0, 1.0), abs=1e-12)
    p, r, f1 = edge_detection_metrics(bad, w, v, u, f, 3, 2)
    assert (p, r) == pytest.approx((2.0 / 3.0, 1.0), abs=1e-12) and f1 == pytest.approx(2.0 / (1.0 / p + 1.0 / r), abs=1e-12)

    from wembed_b200 import cabi
    from wembed_b200.metrics import sample_edge_pairs
    n = 3000
    edges, _, _ = make_problem(n, 4)
    rp, col = cabi.csr_from_edges(n, edges)
    v, u, f = sample_edge_pairs(rp, col, scale=5.0, seed=3)
    m = len(edges)
    assert f[:m].all() and not f[m:].any()                                     # every edge once, first (EdgeSampler.cpp:22-30)
    assert sorted(map(tuple, np.stack([v[:m], u[:m]], 1).tolist())) == sorted(map(tuple, np.sort(edges, 1).tolist()))
    nv, nu = v[m:], u[m:]
    assert (nu > nv).all()                                                     # :52 keeps w > v only
    assert not set(zip(nv.tolist(), nu.tolist())) & set(map(tuple, np.sort(edges, 1).tolist()))
    assert len(set(zip(nv.tolist(), nu.tolist()))) == len(nv)                  # positions are visited once
    assert abs(len(nv) / (5.0 * m) - 1.0) < 0.05                               # expected number = scale * m


@pytest.mark.gpu
def test_device_edge_detection_metric(device_lib):
    """wb_edge_detection against the reference's known answers (tests/TestMetrics.cpp:64-93) and against the
    numpy restatement on a weighted geometric graph after some steps."""
    rp, col, v, u, f, good, bad = _metric_kat_inputs()
    dev = device_lib.DeviceEmbedder(rp, col, embedding_dimension=2)
    dev.set_weights(np.ones(3))
    dev.set_coordinates(good)
    assert dev.edge_detection(v, u, f) == pytest.approx((1.0, 1.0, 1.0), abs=1e-12)
    dev.set_coordinates(bad)
    assert dev.edge_detection(v, u, f) == pytest.approx((2.0 / 3.0, 1.0, 0.8), abs=1e-12)
    assert dev.edge_detection([], [], []) == (-1.0, -1.0, -1.0)

    from wembed_b200.metrics import sample_edge_pairs
    n, d = 4000, 5
    edges, w, x0 = make_problem(n, d)
    rp, col = device_lib.csr_from_edges(n, edges)
    dev = device_lib.DeviceEmbedder(rp, col, embedding_dimension=d, seed=1)
    dev.set_weights(w)
    dev.set_coordinates(x0)
    v, u, f = sample_edge_pairs(rp, col, scale=10.0, seed=2)
    for steps in (1, 60):
        for it in range(1, steps + 1):
            dev.step(lr_exponential(it))
        got = dev.edge_detection(v, u, f)
        exp = edge_detection_metrics(dev.coordinates(), w, v, u, f, n, len(edges))
        assert got == pytest.approx(exp, rel=1e-9), (got, exp)
        assert 0.0 < got[2] <= 1.0
