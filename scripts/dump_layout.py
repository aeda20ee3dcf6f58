"""Dumps the layout (float32 positions) and weights of a bench workload after `steps` steps: input of the offline index experiments."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from wembed_b200 import cabi
name, steps = sys.argv[1], [int(s) for s in sys.argv[2].split(",")]
wl = bench.make_workload(name)
dev = cabi.DeviceEmbedder(wl["row_ptr"], wl["col"], embedding_dimension=wl["d"], seed=1234)
dev.set_weights(wl["weights"]); dev.set_coordinates(wl["x0"])
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
for it in range(1, max(steps) + 1):
    st = dev.step(bench.lr_schedule(it))
    if it in steps:
        np.save(os.path.join(out, f"layout_{name}_{it}.npy"), dev.coordinates().astype(np.float16 if len(steps) > 1 else np.float32))
        print(it, st["num_repulsion_pairs"], st["num_candidates"], st["num_box_tests"], st["rel_displacement"], flush=True)
np.save(os.path.join(out, f"weights_{name}.npy"), wl["weights"].astype(np.float32))
