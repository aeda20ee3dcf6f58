"""Whole-run profile of a bench workload on one GPU: calculateEmbedding's loop (WembedEmbedder.cpp:65-86, loss stop criterion) with
per-phase device times, work counters and the per-step displacement statistics every `every` steps.
usage: gpu_trajectory.py WORKLOAD [MAX_STEPS] [EVERY]      (prints a table; the last line is the time to convergence)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
from helpers import LossMonitor
from wembed_b200 import cabi

name = sys.argv[1]
max_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
every = int(sys.argv[3]) if len(sys.argv) > 3 else 50
wl = bench.make_workload(name)
n, d = wl["n"], wl["d"]
dev = cabi.DeviceEmbedder(wl["row_ptr"], wl["col"], embedding_dimension=d, seed=1234)
dev.set_weights(wl["weights"]); dev.set_coordinates(wl["x0"]); dev.enable_timing(True)
iw = wl["weights"] ** (-1.0 / d)
rho = 1.0 / (iw * iw.max())
mon = LossMonitor()
acc, cnt, dev_ms, t0, builds = {}, 0, 0.0, time.perf_counter(), 0
print("# step | ms/step index repel(search+list) fused recentre | pairs/v pt/v box/v (last step) | rel_disp | disp/rho: mean p99 p999 max | lossA lossR | list")
it = 0
while it < max_steps and not mon.converged():
    it += 1
    probe = it % every == 0
    if probe: xb = dev.coordinates()
    st = dev.step(bench.lr_schedule(it))
    mon.observe(st["loss_attract"] + st["loss_repel"])
    ph = dev.phase_times()
    dev_ms += ph["total"]
    for k, v in ph.items(): acc[k] = acc.get(k, 0.0) + v
    cnt += 1
    builds += int(st["list_rebuilt"])
    if probe:
        dl = np.sqrt(((dev.coordinates() - xb) ** 2).sum(1)) / rho
        q = np.quantile(dl, [0.99, 0.999])
        print(f"{it:5d} | {acc['total']/cnt:7.3f} {acc['index']/cnt:6.3f} {acc['repel']/cnt:7.3f} {acc['attract_update']/cnt:6.3f} {acc['recentre_observe']/cnt:6.3f} | "
              f"{st['num_repulsion_pairs']/n:6.2f} {st['num_candidates']/n:7.1f} {st['num_box_tests']/n:7.1f} | {st['rel_displacement']:.3e} | "
              f"{dl.mean():.4f} {q[0]:.4f} {q[1]:.4f} {dl.max():.4f} | {st['loss_attract']:.5g} {st['loss_repel']:.5g} | "
              f"builds {builds}/{cnt} skin {st['list_skin']:.3f} listed/v {st['num_listed_pairs']/n:.2f} maxratio {st['max_displacement_ratio']:.4f}", flush=True)
        acc, cnt, builds = {}, 0, 0
print(f"converged={mon.converged()} iterations={it} device_ms_total={dev_ms:.1f} wall_s={time.perf_counter()-t0:.2f} (wall includes the probes)")
