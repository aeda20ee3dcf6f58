"""The sharded (multi-GPU) step on ONE GPU: `world` handles of this process joined by wb_comm_init_local (plain device pointers instead
of CUDA IPC mappings) and stepped by wb_step_group, which queues the pieces of the step for all handles in lockstep on one stream (stream
order stands in for the waiting of the barrier kernels).  Everything else a rank does - searching its share of the queries, delivering
pairs to their owners' buffers, the counts matrix, the global sum rows, publishing its rows to the replicas - runs exactly as on `world`
GPUs, and the results must equal the single-handle step bit for bit."""
import numpy as np
import pytest

from helpers import lr_exponential, make_problem

pytestmark = pytest.mark.gpu


def heavy_problem(n, d):
    """Heavy-tailed graph: hub rows (k_hub_rows) and heavy vertices (k_repulse_heavy) take part in the sharded step."""
    from wembed_b200.datasets import degree_weights, heavy_tailed_graph, initial_coordinates
    edges, _ = heavy_tailed_graph(n, 20, seed=3)
    return edges, degree_weights(n, edges, d), initial_coordinates(n, d, seed=5)


@pytest.mark.parametrize("world,n,d,steps,family", [
    (2, 30_000, 4, 6, "geometric"), (8, 60_000, 8, 6, "geometric"), (5, 20_000, 3, 5, "geometric"), (8, 1_000_000, 8, 2, "geometric"),
    (8, 20_000, 8, 4, "heavy"), (8, 30_000, 16, 4, "geometric"), (7, 20_000, 2, 4, "heavy")])
def test_local_group_equals_single_handle(device_lib, monkeypatch, world, n, d, steps, family):
    # room for the dense early steps (~200 partners per vertex for a step or two): a local group cannot grow its buffers
    monkeypatch.setenv("WB_PAIR_CAP", str((40 if n >= 500_000 else 400) * n))
    edges, w, x0 = make_problem(n, d) if family == "geometric" else heavy_problem(n, d)
    rp, col = device_lib.csr_from_edges(n, edges)

    def fresh():
        dev = device_lib.DeviceEmbedder(rp, col, embedding_dimension=d, seed=1234)
        dev.set_weights(w)
        dev.set_coordinates(x0)
        return dev

    single = fresh()
    ref = [single.step(lr_exponential(it)) for it in range(1, steps + 1)]
    x_ref = single.coordinates()
    single.close()
    devs = [fresh() for _ in range(world)]
    device_lib.comm_init_local(devs)
    assert [dv.partition() for dv in devs][0][0] == 0 and devs[-1].partition()[1] == n
    keys = ("loss_attract", "loss_repel", "num_repulsion_pairs", "num_listed_pairs", "sum_displacement", "sum_radius_sq")
    for it in range(1, steps + 1):
        stats = device_lib.step_group(devs, lr_exponential(it))
        for st in stats:
            assert {k: st[k] for k in keys} == {k: ref[it - 1][k] for k in keys}, (it, world)
    for dv in devs:
        assert np.array_equal(dv.coordinates(), x_ref)
        dv.close()
