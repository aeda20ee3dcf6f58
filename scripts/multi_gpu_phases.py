import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from wembed_b200 import cabi, sharding
from helpers import make_problem, lr_exponential
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, d, steps = 1000000, 8, int(sys.argv[1])
edges, w, x0 = make_problem(n, d)
rp, col = cabi.csr_from_edges(n, edges)
dev = cabi.DeviceEmbedder(rp, col, embedding_dimension=d, device=local, seed=1234)
dev.set_weights(w); dev.set_coordinates(x0)
sharding.shard_embedder(dev, rank, world, torch.device("cuda", local))
dev.enable_timing(True)
acc = {}
for it in range(1, steps + 1):
    dev.step(lr_exponential(it))
    if it > steps - 20:
        for k, v in dev.phase_times().items(): acc[k] = acc.get(k, 0) + v / 20
print(f"rank {rank}/{world} mean of last 20 steps:", {k: round(v, 3) for k, v in acc.items()}, flush=True)
dist.barrier(); dist.destroy_process_group()
