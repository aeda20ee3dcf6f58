"""torchrun --nproc-per-node G scripts/multi_gpu_phases.py WORKLOAD WARMUP STEPS: mean phase times of the sharded step
over steps WARMUP+1 .. WARMUP+STEPS of a bench workload (per rank; the step time is the max over ranks)."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
import bench
from wembed_b200 import cabi, sharding
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
name, warm, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
wl = bench.make_workload(name, rank, world)
dev = cabi.DeviceEmbedder(wl["row_ptr"], wl["col"], embedding_dimension=wl["d"], device=local, seed=1234)
dev.set_weights(wl["weights"]); dev.set_coordinates(wl["x0"])
sharding.shard_embedder(dev, rank, world, torch.device("cuda", local))
dev.enable_timing(True)
acc = {}
for it in range(1, warm + steps + 1):
    st = dev.step(bench.lr_schedule(it))
    if it > warm:
        for k, v in dev.phase_times().items(): acc[k] = acc.get(k, 0) + v / steps
tot = torch.tensor([acc["total"]], device="cuda", dtype=torch.float64)
dist.all_reduce(tot, op=dist.ReduceOp.MAX)
print(f"rank {rank}/{world} {name} steps {warm+1}..{warm+steps}:", {k: round(v, 3) for k, v in acc.items()},
      f"| max-over-ranks total {tot.item():.3f} ms -> {1e3/tot.item():.2f} steps/s | pairs {st['num_repulsion_pairs']:.0f} lossA {st['loss_attract']:.8g}", flush=True)
dist.barrier(); dist.destroy_process_group()
