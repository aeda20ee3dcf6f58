"""torchrun --nproc-per-node G scripts/multi_gpu_check.py : sharded step vs the single-GPU step on the same problem."""
import os, sys, time
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from wembed_b200 import cabi, sharding
from helpers import make_problem, lr_exponential

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, d, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
if len(sys.argv) > 4 and sys.argv[4] == "heavy":      # heavy-tailed graph: exercises k_attract_hubs and k_repulse_heavy in the sharded step
    from wembed_b200.datasets import degree_weights, heavy_tailed_graph, initial_coordinates
    edges, _ = heavy_tailed_graph(n, 20, seed=3)
    w, x0 = degree_weights(n, edges, d), initial_coordinates(n, d, seed=5)
else:
    edges, w, x0 = make_problem(n, d)
rp, col = cabi.csr_from_edges(n, edges)

def run(shard):
    dev = cabi.DeviceEmbedder(rp, col, embedding_dimension=d, device=local, seed=1234)
    dev.set_weights(w); dev.set_coordinates(x0)
    if shard:
        own = sharding.shard_embedder(dev, rank, world, torch.device("cuda", local))
    stats = []
    torch.cuda.synchronize(); dist.barrier(); t0 = time.time()
    warm = min(10, steps // 2)           # the first collectives set up NCCL channels: keep them out of the timing
    for it in range(1, steps + 1):
        if it == warm + 1:
            dev.mark(0)
        stats.append(dev.step(lr_exponential(it)))
        if os.environ.get("WB_DEBUG") and (it <= 12 or it % 10 == 0):
            print(f"[check rank {rank}] shard={shard} step {it} pairs {stats[-1]['num_repulsion_pairs']:.0f} listed {stats[-1]['num_listed_pairs']:.0f}", file=sys.stderr, flush=True)
    dev.mark(1)
    ms = dev.elapsed_ms(0, 1) * steps / (steps - warm)
    return dev.coordinates(), stats, ms

xs, ss, ms_s = run(True)
x1, s1, ms_1 = run(False)
err = np.abs(xs - x1).max() / max(1.0, np.abs(x1).max())
keys = ("loss_attract", "loss_repel", "num_repulsion_pairs", "rel_displacement")
worst = max(abs(a[k] - b[k]) / max(1e-30, abs(b[k])) for a, b in zip(ss, s1) for k in keys)
allx = [torch.zeros(xs.shape, dtype=torch.float64, device="cuda") for _ in range(world)]
dist.all_gather(allx, torch.tensor(xs, device="cuda"))
same = all(bool((a == allx[0]).all()) for a in allx)
if rank == 0:
    print(f"world={world} n={n} d={d} steps={steps}: sharded vs single max rel coord err {err:.3e}, worst stat rel err {worst:.3e}, "
          f"pairs {ss[-1]['num_repulsion_pairs']:.0f} vs {s1[-1]['num_repulsion_pairs']:.0f}, replicas identical across ranks: {same}, "
          f"ms/step sharded {ms_s/steps:.3f} single {ms_1/steps:.3f} speedup {ms_1/ms_s:.2f}x", flush=True)
    assert same and err < 1e-5 and worst < 1e-6, (same, err, worst)
dist.barrier()
dist.destroy_process_group()
