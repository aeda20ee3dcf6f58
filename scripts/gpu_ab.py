"""A/B timing of repulsion variants along the c3 trajectory (prints phase times)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from wembed_b200 import cabi
from helpers import make_problem, lr_exponential
n, d, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
edges, w, x0 = make_problem(n, d)
rp, col = cabi.csr_from_edges(n, edges)
dev = cabi.DeviceEmbedder(rp, col, embedding_dimension=d, seed=1234)
dev.set_weights(w); dev.set_coordinates(x0); dev.enable_timing(True)
tot = 0.0
for it in range(1, steps + 1):
    st = dev.step(lr_exponential(it)); ph = dev.phase_times(); tot += ph['total']
    if it <= 5 or it % 10 == 0:
        print(f"var={os.environ.get('WB_REPULSE_VARIANT','1')} n={n} d={d} it={it} pairs/v {st['num_repulsion_pairs']/n:.2f} tests/v {st['num_candidates']/n:.0f} box/v {st['num_box_tests']/n:.0f} lossA {st['loss_attract']:.6g} lossR {st['loss_repel']:.6g} | index {ph['index']:.3f} attract {ph['attract_update']:.3f} repel {ph['repel']:.3f} recentre {ph['recentre_observe']:.3f} total {ph['total']:.3f} ms", flush=True)
print(f"sum of step times over {steps} steps: {tot:.1f} ms")
