"""Runs `steps` steps of a workload (for ncu captures; prints nothing that is a bench value)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from wembed_b200 import cabi
from helpers import make_problem, lr_exponential
n, d, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
edges, w, x0 = make_problem(n, d)
rp, col = cabi.csr_from_edges(n, edges)
dev = cabi.DeviceEmbedder(rp, col, embedding_dimension=d, seed=1234)
dev.set_weights(w); dev.set_coordinates(x0)
for it in range(1, steps + 1):
    st = dev.step(lr_exponential(it))
print("done", st["iteration"], st["num_repulsion_pairs"])
