"""c5-shaped run (geometric, avg degree 20, d=16) at a given n: memory + time per step."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wembed_b200 import cabi
from wembed_b200.datasets import degree_weights, geometric_graph, initial_coordinates
n, steps = int(sys.argv[1]), int(sys.argv[2]); d = int(sys.argv[3]) if len(sys.argv) > 3 else 16
t = time.time(); edges, _ = geometric_graph(n, 20, 42); print("generated", len(edges), f"{time.time()-t:.1f}s", flush=True)
w, x0 = degree_weights(n, edges, d), initial_coordinates(n, d, seed=1234)
t = time.time(); rp, col = cabi.csr_from_edges(n, edges); print("csr", f"{time.time()-t:.1f}s", flush=True)
del edges
dev = cabi.DeviceEmbedder(rp, col, embedding_dimension=d, seed=1234)
dev.set_weights(w); dev.set_coordinates(x0); dev.enable_timing(True)
import torch; print("device memory used GB", (torch.cuda.mem_get_info()[1]-torch.cuda.mem_get_info()[0])/1e9, flush=True)
lr = lambda it: 10 * 0.995 ** it * (it / 20 if it < 20 else 1)
for it in range(1, steps + 1):
    st = dev.step(lr(it)); ph = dev.phase_times()
    if it <= 6 or it % 10 == 0:
        print(f"c5 n={n} d={d} it={it} pairs/v {st['num_repulsion_pairs']/n:.2f} pt/v {st['num_candidates']/n:.0f} box/v {st['num_box_tests']/n:.0f} lossA {st['loss_attract']:.5g} lossR {st['loss_repel']:.5g} | index {ph['index']:.3f} attract {ph['attract_update']:.3f} repel {ph['repel']:.3f} recentre {ph['recentre_observe']:.3f} total {ph['total']:.3f} ms", flush=True)
