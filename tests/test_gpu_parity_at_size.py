"""Parity at BASELINE.json's sizes: the CUDA path (C ABI) against the CPU oracle on the configurations themselves.

  c2  geometric n = 1e5, d = 4      3 steps against the reference's own code (oracle/_ref, WembedEmbedder.cpp:13-63) and the port
  c3  geometric n = 1e6, d = 8      one step from the device's step-30 and step-100 layouts against the port
  c4  heavy-tailed n = 1e6, d = 8   one step against the port with the heavy-vertex walk active

and, because the port's index is a box hierarchy like the device's, an index-free census of the repulsive pairs
(helpers.brute_force_repulsive_pairs: all n^2 pairs in tiles with plain torch) at c3 and c4.
Tolerances are those of tests/test_gpu_parity.py (north_star: CSR / pair sets exact, forces and coordinates 1e-5 relative,
near-hinge vertices masked and counted).
"""
import numpy as np
import pytest

import oracle
from helpers import brute_force_repulsive_pairs, lr_exponential, make_problem, near_threshold_edge_owners
from test_gpu_parity import COORD_RTOL, StepParity, assert_step_close

pytestmark = pytest.mark.gpu


def _one_step_against_port(device_lib, edges, n, d, w, rp, col, x, max_flagged_frac=0.02, coord_rtol=COORD_RTOL, pair_slack=8, tile=2048):
    """Fresh handles on both sides (zero optimizer state), same layout, same iteration counter: one step each.  The near-hinge
    mask and the expected number of repulsive pairs come from the index-free census, not from either implementation."""
    lo, hi, deg_lo, deg_hi = brute_force_repulsive_pairs(x, w, rp, col, tile=tile)
    flagged = (deg_lo != deg_hi) | near_threshold_edge_owners(x, w, rp, col)
    cpu = oracle.CpuEmbedder("port", edges, n=n, embeddingDimension=d, init_state=False)
    np.testing.assert_array_equal(cpu.csr()[0], rp)
    np.testing.assert_array_equal(cpu.csr()[1], col)
    dev = device_lib.DeviceEmbedder(rp, col, embedding_dimension=d, keep_forces=1, seed=1234)
    for e in (cpu, dev):
        e.set_weights(w)
        e.set_coordinates(x)
    cpu.step()                                              # iteration 1 on both sides: lr of step 1, Adam t = 1
    st = dev.step(lr_exponential(1))
    cs = cpu.stats()
    tracker = StepParity(w, rp, col, d, max_flagged_frac)
    tracker.coord_rtol = coord_rtol
    assert_step_close(None, w, rp, col, cpu.forces(), dev.forces(), cpu.coordinates(), dev.coordinates(), max_flagged_frac=max_flagged_frac,
                      flagged=flagged, tracker=tracker)
    assert abs(st["num_repulsion_pairs"] - cs["num_rep_pairs"]) <= pair_slack, (st["num_repulsion_pairs"], cs["num_rep_pairs"])
    np.testing.assert_allclose(st["loss_attract"], cs["loss_attract"], rtol=1e-5)
    np.testing.assert_allclose(st["loss_repel"], cs["loss_repel"], rtol=1e-4, atol=1e-6)
    assert lo <= st["num_repulsion_pairs"] <= hi, (lo, st["num_repulsion_pairs"], hi)
    assert lo <= cs["num_rep_pairs"] <= hi, (lo, cs["num_rep_pairs"], hi)
    assert hi - lo <= 4e-4 * max(hi, 1) + 8             # the shell |dist ws - L| <= 1e-5 L holds ~2 * 1e-5 * d of all pairs
    out = dict(pairs=st["num_repulsion_pairs"], lo=lo, hi=hi, flagged=int(flagged.sum()), unstable=float(tracker.unstable.mean()),
               force_err=float(np.abs(cpu.forces() - dev.forces())[~flagged].max() / np.abs(cpu.forces()).max()))
    cpu.close()
    dev.close()
    return out


def test_c2_three_steps_against_the_reference(device_lib, port_lib, ref_lib):
    """BASELINE.json configs[1]: n = 1e5, d = 4, "per-step parity vs reference CPU"."""
    n, d = 100_000, 4
    edges, w, x0 = make_problem(n, d)
    rp, col = device_lib.csr_from_edges(n, edges)
    ref = oracle.CpuEmbedder("ref", edges, n=n, embeddingDimension=d, init_state=False)
    cpu = oracle.CpuEmbedder("port", edges, n=n, embeddingDimension=d, init_state=False)
    for e in (ref, cpu):                                    # CSR indexing bit-exact against the reference's Graph
        np.testing.assert_array_equal(e.csr()[0], rp)
        np.testing.assert_array_equal(e.csr()[1], col)
    dev = device_lib.DeviceEmbedder(rp, col, embedding_dimension=d, keep_forces=1, seed=1234)
    for e in (ref, cpu, dev):
        e.set_weights(w)
        e.set_coordinates(x0)
    tracker = None
    for it in range(1, 4):
        flagged = cpu.near_threshold(1e-5)
        ref.step()
        cpu.step()
        st = dev.step(lr_exponential(it))
        rs, cs = ref.stats(), cpu.stats()
        assert st["iteration"] == it == rs["iteration"]
        # the port against the reference itself at this size (fp64 both: summation order only)
        assert np.abs(ref.forces() - cpu.forces()).max() <= 1e-9 * np.abs(ref.forces()).max()
        tracker = assert_step_close(None, w, rp, col, ref.forces(), dev.forces(), ref.coordinates(), dev.coordinates(), flagged=flagged, tracker=tracker)
        # the reference does not export its pair counter; its restatement does, and the two agree on every force above
        assert abs(st["num_repulsion_pairs"] - cs["num_rep_pairs"]) <= 8
        np.testing.assert_allclose(st["loss_attract"], rs["loss_attract"], rtol=1e-5)
        np.testing.assert_allclose(st["loss_repel"], rs["loss_repel"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(st["rel_displacement"], rs["rel_displacement"], rtol=1e-3)
        for e in (ref, cpu):
            e.set_coordinates(dev.coordinates())            # stay on one trajectory


@pytest.mark.parametrize("steps", [30, 100])
def test_c3_one_step_against_the_port_and_a_brute_force_census(device_lib, port_lib, steps):
    """BASELINE.json configs[2] (the headline): n = 1e6, d = 8, from the layouts the trajectory itself produces."""
    n, d = 1_000_000, 8
    edges, w, x0 = make_problem(n, d)
    rp, col = device_lib.csr_from_edges(n, edges)
    walker = device_lib.DeviceEmbedder(rp, col, embedding_dimension=d, seed=1234)
    walker.set_weights(w)
    walker.set_coordinates(x0)
    for it in range(1, steps + 1):
        walker.step(lr_exponential(it))
    x = walker.coordinates()
    walker.close()
    res = _one_step_against_port(device_lib, edges, n, d, w, rp, col, x)
    lo, hi = res["lo"], res["hi"]
    if steps >= 100:
        assert res["pairs"] > n                             # a dense state: more than one repulsive partner per vertex
    print(f"c3 step {steps}: pairs {res['pairs']:.0f} in [{lo}, {hi}], near-hinge vertices {res['flagged']}, max force err {res['force_err']:.2e}")


def test_c4_one_step_against_the_port_with_heavy_vertices(device_lib, port_lib):
    """BASELINE.json configs[3]: heavy-tailed n = 1e6, average degree 20, d = 8 (hubs: k_repulse_heavy, k_attract_hubs)."""
    from wembed_b200.datasets import degree_weights, heavy_tailed_graph, initial_coordinates
    n, d, steps = 1_000_000, 8, 30
    edges, _ = heavy_tailed_graph(n, 20, seed=42)
    w, x0 = degree_weights(n, edges, d), initial_coordinates(n, d, seed=1234)
    rp, col = device_lib.csr_from_edges(n, edges)
    assert (w >= 32.0 * w.mean()).sum() > 100                # heavy vertices exist
    assert np.diff(rp).max() > 10_000
    walker = device_lib.DeviceEmbedder(rp, col, embedding_dimension=d, seed=1234)
    walker.set_weights(w)
    walker.set_coordinates(x0)
    for it in range(1, steps + 1):
        walker.step(lr_exponential(it))
    x = walker.coordinates()
    walker.close()
    # hub rows sum 1e4..1e5 fp32 terms: same coordinate tolerance as the heavy-tailed test at n = 2e4
    res = _one_step_against_port(device_lib, edges, n, d, w, rp, col, x, max_flagged_frac=0.05, coord_rtol=5e-5, pair_slack=64, tile=1024)
    lo, hi = res["lo"], res["hi"]
    assert res["pairs"] > 10_000
    print(f"c4 step {steps}: pairs {res['pairs']:.0f} in [{lo}, {hi}], near-hinge vertices {res['flagged']}, max force err {res['force_err']:.2e}")
