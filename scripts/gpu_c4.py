"""c4 (heavy-tailed, n=1M, avg degree 20, d=8) trajectory timing."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wembed_b200 import cabi
from wembed_b200.datasets import degree_weights, heavy_tailed_graph, initial_coordinates
n, d, steps = int(sys.argv[1]), 8, int(sys.argv[2])
t = time.time(); edges, _ = heavy_tailed_graph(n, 20); print("generated", len(edges), time.time() - t, flush=True)
w, x0 = degree_weights(n, edges, d), initial_coordinates(n, d, seed=1234)
rp, col = cabi.csr_from_edges(n, edges)
print("max degree", np.diff(rp).max(), "w range", w.min(), w.max(), flush=True)
dev = cabi.DeviceEmbedder(rp, col, embedding_dimension=d, seed=1234)
dev.set_weights(w); dev.set_coordinates(x0); dev.enable_timing(True)
lr = lambda it: 10 * 0.995 ** it * (it / 20 if it < 20 else 1)
for it in range(1, steps + 1):
    st = dev.step(lr(it)); ph = dev.phase_times()
    if it <= 5 or it % 10 == 0:
        print(f"c4 n={n} it={it} pairs/v {st['num_repulsion_pairs']/n:.2f} tests/v {st['num_candidates']/n:.0f} box/v {st['num_box_tests']/n:.0f} lossA {st['loss_attract']:.5g} lossR {st['loss_repel']:.5g} | index {ph['index']:.3f} attract {ph['attract_update']:.3f} repel {ph['repel']:.3f} recentre {ph['recentre_observe']:.3f} total {ph['total']:.3f} ms", flush=True)
