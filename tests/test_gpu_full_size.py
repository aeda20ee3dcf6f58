"""BASELINE.json's full sizes, checked through size-independent properties (the oracle cannot run here in seconds):
pairwise forces are antisymmetric so they sum to zero, every unordered repulsive pair is seen from both sides, the
recentred layout has zero mean, and the whole trajectory is bit-reproducible."""
import numpy as np
import pytest

from helpers import lr_exponential, make_problem

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,d,steps", [(100_000, 4, 12), (1_000_000, 8, 8)])
def test_properties_at_baseline_sizes(device_lib, n, d, steps):
    edges, w, x0 = make_problem(n, d)
    rp, col = device_lib.csr_from_edges(n, edges)
    runs = []
    for _ in range(2):
        dev = device_lib.DeviceEmbedder(rp, col, embedding_dimension=d, keep_forces=1, seed=1234)
        dev.set_weights(w)
        dev.set_coordinates(x0)
        stats = [dev.step(lr_exponential(it)) for it in range(1, steps + 1)]
        runs.append((dev.coordinates(), dev.forces(), stats))
    x, f, stats = runs[0]
    assert np.array_equal(x, runs[1][0]) and np.array_equal(f, runs[1][1])
    for a, b in zip(stats, runs[1][2]):
        assert a == b
    for s in stats:
        assert s["num_repulsion_pairs"] % 2 == 0
        assert np.isfinite([s["loss_attract"], s["loss_repel"], s["rel_displacement"]]).all()
    # Newton's third law: attraction and repulsion are evaluated once from each side with the same magnitude
    assert np.abs(f.sum(axis=0)).max() <= 1e-4 * np.abs(f).sum(axis=0).max()
    # applyGravityCentre: the stored layout is centred
    assert np.abs(x.mean(axis=0)).max() <= 1e-4 * np.abs(x).max()
    # the loss of the attractive hinge is reproduced from the returned layout of the previous step
    assert stats[0]["loss_attract"] > 0 and stats[-1]["iteration"] == steps
