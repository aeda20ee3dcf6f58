"""Known-answer tests taken from the reference's own unit tests, driven through the CPU port (and, with -m gpu,
through the C ABI in test_gpu_*.py).  Source of every expected value is cited."""
import numpy as np
import pytest

import oracle
from wembed_b200 import cabi


def rows(rp, col):
    return [list(col[rp[v]:rp[v + 1]]) for v in range(len(rp) - 1)]


def test_graph_from_edge_list(port_lib):
    """tests/TestGraph.cpp:61-139: CSR from an edge list - symmetric, deduplicated, neighbours ascending."""
    edges = [(0, 1), (0, 2), (1, 2), (2, 3), (3, 0), (0, 1), (1, 0)]   # repeated + reversed duplicates
    expect = [[1, 2, 3], [0, 2], [0, 1, 3], [0, 2]]
    cpu = oracle.CpuEmbedder("port", edges, init_state=False)
    assert rows(*cpu.csr()) == expect
    assert rows(*cabi.csr_from_edges(4, edges)) == expect
    truth = {(0, 1): True, (1, 0): True, (0, 3): True, (1, 3): False, (3, 1): False, (2, 3): True, (0, 0): False}
    for (a, b), t in truth.items():
        assert cpu.are_neighbors(a, b) == t


def test_graph_self_loop_dropped(port_lib):
    """tests/TestGraph.cpp:22-29 / Graph.cpp:124-128: a self loop is ignored (we drop all of them)."""
    edges = [(0, 1), (1, 1), (1, 2), (2, 2)]
    cpu = oracle.CpuEmbedder("port", edges, init_state=False)
    assert rows(*cpu.csr()) == [[1], [0, 2], [1]]
    assert rows(*cabi.csr_from_edges(3, edges)) == [[1], [0, 2], [1]]


def test_isolated_and_missing_ids(port_lib):
    """Graph.cpp:101-106: n = largest id + 1; ids without edges become isolated vertices."""
    edges = [(0, 4), (4, 2)]
    cpu = oracle.CpuEmbedder("port", edges, init_state=False)
    assert cpu.n == 5
    assert rows(*cpu.csr()) == [[4], [], [4], [], [0, 2]]


# tests/TestSNNQueries.cpp:6-64 - the only known-answer test of a radius query in the reference.
SNN_POINTS = {0: (0.5, 2.0, 3.0), 2: (0.0, 1.0, 3.0), 1: (0.0, 1.0, 0.0)}
SNN_QUERIES = [((0.0, 1.5, 3.1), 1.0, {0, 2}), ((0.0, 1.5, 3.1), 0.5, set()), ((0.5, 0.0, 4.0), 1.7, {2})]


def snn_case(query, radius):
    """The KAT as an embedding problem: unit weights, edgeLength = radius, the query is an extra vertex 3."""
    x = np.zeros((4, 3))
    for i, p in SNN_POINTS.items():
        x[i] = p
    x[3] = query
    edges = [(0, 1), (1, 2), (2, 3)]   # any connected graph; candidate sets ignore adjacency
    return x, edges


@pytest.mark.parametrize("query,radius,expect", SNN_QUERIES)
def test_snn_radius_query_kat_port(port_lib, query, radius, expect):
    x, edges = snn_case(query, radius)
    cpu = oracle.CpuEmbedder("port", edges, n=4, init_state=False, embeddingDimension=3, edgeLength=radius)
    cpu.set_weights(np.ones(4))
    cpu.set_coordinates(x)
    assert set(cpu.candidates(3).tolist()) - {3} == expect


def test_snn_first_query_single_point(port_lib):
    """TestSNNQueries.cpp:9-21: one point, query at the point with r = 0.1 returns it."""
    cpu = oracle.CpuEmbedder("port", [(0, 1)], n=2, init_state=False, embeddingDimension=3, edgeLength=0.1)
    cpu.set_weights(np.ones(2))
    cpu.set_coordinates(np.array([[0.5, 2.0, 3.0], [50.0, 50.0, 50.0]]))
    assert cpu.candidates(0).tolist() == [0]
