"""The CPU restatement (oracle/wembed_port.cpp) against golden vectors generated from the reference's own sources
(tests/golden/make_golden.py), and - where the reference checkout exists - against the reference itself."""
import os

import numpy as np
import pytest

import oracle
from helpers import SMALL_GRAPH, ring_graph

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name))


def replay(cpu, g, prefix="", resync=True, rtol=1e-9):
    """Steps `cpu` along the golden trace; coordinates are re-synchronised to the golden state after every step so
    summation-order rounding cannot accumulate through the chaotic dynamics."""
    X, F, S = g[prefix + "x"], g[prefix + "f"], g[prefix + "stats"]
    for i in range(len(X)):
        cpu.step()
        s = cpu.stats()
        scale = max(1.0, np.abs(F[i]).max())
        assert np.abs(cpu.forces() - F[i]).max() <= rtol * scale, f"forces, step {i}"
        assert np.abs(cpu.coordinates() - X[i]).max() <= rtol * max(1.0, np.abs(X[i]).max()), f"coordinates, step {i}"
        got = [s["loss_attract"], s["loss_repel"], s["lr"], s["rel_displacement"]]
        np.testing.assert_allclose(got, S[i][:4], rtol=1e-9, atol=1e-12, err_msg=f"stats, step {i}")
        assert s["iteration"] == S[i][5]
        if np.isfinite(S[i][4]):
            np.testing.assert_allclose(s["rel_loss_improvement"], S[i][4], rtol=1e-6, atol=1e-12)
        if resync:
            cpu.set_coordinates(X[i])


@pytest.mark.parametrize("kind", ["port", "ref"])
def test_ring64_coincident_start(kind, request):
    """tests/TestDeterminism.cpp:96-109 protocol: all nodes coincident -> the dist <= 0 random-direction branch
    (Rand::localGenerator + setToRandomUnitVector) is exercised by every pair of step 1."""
    request.getfixturevalue(f"{kind}_lib")
    g = load("ring64_d2_coincident.npz")
    cpu = oracle.CpuEmbedder(kind, g["edges"], seed=1234, embeddingDimension=2, maxIterations=1000)
    np.testing.assert_array_equal(cpu.weights(), g["w"])
    cpu.set_coordinates(np.zeros((64, 2)))
    replay(cpu, g, resync=False, rtol=1e-7)


@pytest.mark.parametrize("kind", ["port", "ref"])
def test_ring64_random_layout_matches_reference_rng(kind, request):
    """Initial layout = Rand::randomCoordinates from mt19937(1234) (Rand.cpp:101-109); must be bit-identical."""
    request.getfixturevalue(f"{kind}_lib")
    g = load("ring64_d2_random.npz")
    cpu = oracle.CpuEmbedder(kind, g["edges"], seed=1234, embeddingDimension=2, maxIterations=1000)
    np.testing.assert_array_equal(cpu.coordinates(), g["x0"])
    replay(cpu, g)


@pytest.mark.parametrize("seed", [1, 2])
def test_small_graph_converges_like_the_reference(port_lib, seed):
    """BASELINE.json configs[0]: assets/small_graph.edg, default options; degree weights and the final loss are pinned;
    the stop iteration is chaotic in the last bits of the sums, so it is pinned to a band around the reference's."""
    g = load(f"small_graph_seed{seed}.npz")
    cpu = oracle.CpuEmbedder("port", SMALL_GRAPH, seed=seed)
    np.testing.assert_array_equal(cpu.coordinates(), g["x0"])
    np.testing.assert_allclose(cpu.weights(), g["w"], rtol=1e-15)
    np.testing.assert_allclose(cpu.weights(), [0.41666667, 1.25, 1.25, 1.25, 0.83333333], rtol=1e-7)
    rp, col = cpu.csr()
    np.testing.assert_array_equal(rp, g["csr_row"])
    np.testing.assert_array_equal(col, g["csr_col"])
    iters = cpu.run()
    s = cpu.stats()
    assert s["loss_attract"] + s["loss_repel"] == float(g["loss_final"]) == 0.0
    assert 0.6 * int(g["iterations"]) <= iters <= 1.6 * int(g["iterations"])


@pytest.mark.parametrize("d", [4, 8])
def test_geometric_trace_and_candidates(port_lib, d):
    g = load(f"geo600_d{d}.npz")
    n = len(g["w"])
    cpu = oracle.CpuEmbedder("port", g["edges"], n=n, seed=99, embeddingDimension=d)
    np.testing.assert_allclose(cpu.weights(), g["w"], rtol=1e-15)
    rp, col = cpu.csr()
    np.testing.assert_array_equal(rp, g["csr_row"])
    np.testing.assert_array_equal(col, g["csr_col"])
    cpu.set_coordinates(g["x0"])
    sub = {k: g[k][:3] for k in ("x", "f", "stats")}
    replay(cpu, sub)
    # candidate sets at the state after step 3: identical id sets (SNN semantics: all points within the class radius)
    offs, ids = g["cand_offsets"], g["cand_ids"]
    for i, q in enumerate(g["queries"]):
        np.testing.assert_array_equal(cpu.candidates(int(q)), ids[offs[i]:offs[i + 1]])
    sub = {k: g[k][3:] for k in ("x", "f", "stats")}
    replay(cpu, sub)


VARIANTS = {
    "simple": dict(optimizerType=0, simpleOptMaxDisplacement=0.5),
    "centre": dict(centreScale=0.05),
    "unit": dict(weightType=0),
    "hint": dict(dimensionHint=2.0, embeddingDimension=3),
    "scales": dict(attractionScale=2.0, repulsionScale=0.5, edgeLength=1.5),
    "adaptive": dict(lrScheduleType=1, lossRateWindow=3, lrAdaptPatience=2),
}


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_option_variants(port_lib, name):
    g = load("geo300_options.npz")
    cpu = oracle.CpuEmbedder("port", g["edges"], n=300, seed=7, **VARIANTS[name])
    np.testing.assert_allclose(cpu.weights(), g[f"{name}_w"], rtol=1e-15)
    cpu.set_coordinates(g[f"{name}_x0"])
    replay(cpu, g, prefix=name + "_")


def test_port_matches_reference_live(ref_lib, port_lib):
    """Where the reference checkout exists: a larger case than the fixtures, straight against the reference."""
    from wembed_b200.datasets import geometric_graph
    n, d = 3000, 8
    edges, _ = geometric_graph(n, 10, seed=21)
    r = oracle.CpuEmbedder("ref", edges, n=n, seed=3, embeddingDimension=d)
    p = oracle.CpuEmbedder("port", edges, n=n, seed=3, embeddingDimension=d)
    np.testing.assert_array_equal(r.coordinates(), p.coordinates())
    for a, b in zip(r.csr(), p.csr()):
        np.testing.assert_array_equal(a, b)
    for _ in range(5):
        r.step()
        p.step()
        fr, fp = r.forces(), p.forces()
        assert np.abs(fr - fp).max() <= 1e-12 * np.abs(fr).max()
        assert np.abs(r.coordinates() - p.coordinates()).max() <= 1e-10
        p.set_coordinates(r.coordinates())
    for q in (0, 17, 1234, 2999):
        np.testing.assert_array_equal(np.sort(r.candidates(q)), p.candidates(q))
    assert ring_graph(8).shape == (16, 2)
