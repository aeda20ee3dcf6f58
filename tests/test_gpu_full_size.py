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


def test_box_formats_agree_at_the_headline_size(device_lib, monkeypatch):
    """c3's size (n = 1e6, d = 8): the walk over the half-precision boxes and the walk over the fp32 boxes find the same pairs,
    so layouts and counters agree bit for bit (the repulsion rows are exact integer sums)."""
    n, d, steps = 1_000_000, 8, 6
    edges, w, x0 = make_problem(n, d)
    rp, col = device_lib.csr_from_edges(n, edges)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("WB_HALF_BOXES", mode)
        dev = device_lib.DeviceEmbedder(rp, col, embedding_dimension=d, seed=1234)
        dev.set_weights(w)
        dev.set_coordinates(x0)
        stats = [dev.step(lr_exponential(it)) for it in range(1, steps + 1)]
        out[mode] = (dev.coordinates(), [(s["num_repulsion_pairs"], s["loss_repel"], s["loss_attract"]) for s in stats],
                     [s["num_box_tests"] for s in stats])
        dev.close()
    assert out["0"][1] == out["1"][1]
    assert np.array_equal(out["0"][0], out["1"][0])
    assert out["0"][1][-1][0] > 0
    # the outward rounding may only ADD box tests, and only a few
    for a, b in zip(out["0"][2], out["1"][2]):
        assert a <= b <= 1.1 * a
