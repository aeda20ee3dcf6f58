"""How much of the fused step kernel's time is the randomness of the partner gathers?  Runs a bench workload twice: with the
generator's vertex numbering (random with respect to the geometry) and renumbered along a 2-D Morton curve of the generator's
points (the best locality any graph-based renumbering could reach).  Prints mean phase times over steps w+1..w+k.
usage: gpu_locality_probe.py [n] [d] [warmup] [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from wembed_b200 import cabi
from wembed_b200.datasets import degree_weights, geometric_graph, initial_coordinates

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 20
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 60
edges, pts = geometric_graph(n, 10, 42)
x0 = initial_coordinates(n, d, seed=1234)


def morton2(p):
    q = np.floor(p / p.max() * 65535).astype(np.uint64)
    def spread(v):
        v = (v | (v << 8)) & 0x00FF00FF
        v = (v | (v << 4)) & 0x0F0F0F0F
        v = (v | (v << 2)) & 0x33333333
        v = (v | (v << 1)) & 0x55555555
        return v
    return spread(q[:, 0]) | (spread(q[:, 1]) << 1)


def run(tag, e, x):
    rp, col = cabi.csr_from_edges(n, e)
    w = degree_weights(n, e, d)
    dev = cabi.DeviceEmbedder(rp, col, embedding_dimension=d, seed=1234)
    dev.set_weights(w); dev.set_coordinates(x); dev.enable_timing(True)
    acc = {}
    for it in range(1, warm + steps + 1):
        st = dev.step(bench.lr_schedule(it))
        if it > warm:
            for k, v in dev.phase_times().items(): acc[k] = acc.get(k, 0.0) + v / steps
    print(tag, " ".join(f"{k} {v:.4f}" for k, v in acc.items()), f"| pairs {st['num_repulsion_pairs']:.0f} lossA {st['loss_attract']:.8g}", flush=True)
    dev.close()


run("generator order", edges, x0)
order = np.argsort(morton2(pts), kind="stable")          # new id k = old vertex order[k]
new_id = np.empty(n, np.int64); new_id[order] = np.arange(n)
run("morton order   ", new_id[edges].astype(np.int32), x0[order])
