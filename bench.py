#!/usr/bin/env python
"""bench.py - WEmbed gradient-descent step throughput on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|c5|small]

One "step" = one WembedEmbedder::calculateStep over the whole graph (index rebuild, attractive and
repulsive forces, Adam, recentring, observables).  Metric = directed edge-force updates per second =
2m * steps / seconds (BASELINE.json "edge-force updates/sec"); steps/s is reported beside it.

Workload (N = 1): BASELINE.json configs[2] - synthetic 2-D geometric random graph, n = 1M, average degree 10,
embedded in d = 8 with default options (Adam, ExponentialCooling lr 10, L = 1), trajectory started from the
reference's initial layout (uniform cube) with W warm-up steps then K timed steps.

Lines printed (one JSON object on stdout, rank 0):
  value     device-resident throughput: K asynchronous steps, CUDA events on the handle's stream
  e2e       through the blocking C ABI with host buffers: wb_set_coordinates(host) + K x wb_step (per-step
            D2H of the observables) + wb_get_coordinates(host), all inside the timed region
  roofline  dominant kernel (repulsion search) and, per kernel, algorithmic bytes / measured time vs the measured HBM peak
  cpu_baseline  oracle port (oracle/wembed_port.cpp, OpenMP): 3 timed steps from the same state, + parity_at_config
  same_sample / convergence / secondary (c5) / ranks_identical (N > 1): see DESIGN.md section 6
--impl reference times the reference's own C++ (oracle/_ref, SNN index) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n, avg_degree, d, family)
    "small": (100_000, 10, 8, "geometric"),     # the reference arm's bounded sample (its SNN index cannot finish a step at n = 1e6)
    "c2": (100_000, 10, 4, "geometric"),
    "c3": (1_000_000, 10, 8, "geometric"),
    "c4": (1_000_000, 20, 8, "heavy_tailed"),
    "c5": (10_000_000, 20, 16, "geometric"),      # ~1e8 undirected edges
}


def lr_schedule(it, lr0=10.0, cooling=0.995, warmup=20):
    """ExponentialCoolingSchedule + warm-up (LRScheduler.cpp:7-17), iterations are 1-based."""
    lr = lr0 * cooling ** float(it)
    return lr * it / warmup if it < warmup else lr


def make_workload(name, rank=0, world=1):
    """Synthetic graph + the reference's initial state.  The CSR comes out of the PRODUCT's ingestion path (wembed::graphFromEdges,
    wembed_b200/host/graph.cpp - what replaces the reference's std::map<int, std::set<int>> build, Graph.cpp:87-150) and is timed."""
    from wembed_b200 import host
    from wembed_b200.datasets import degree_weights, geometric_graph, heavy_tailed_graph, initial_coordinates
    n, deg, d, family = WORKLOADS[name]
    seed = 42         # every rank builds the same graph: at N > 1 it is sharded by vertex range (strong scaling)
    cache = os.path.join("/tmp", f"wembed_wl_{name}_{seed}.npy")
    if world > 1 and rank != 0:            # one generator per box: the other ranks wait for rank 0's file
        while not os.path.exists(cache):
            time.sleep(0.2)
    if os.path.exists(cache):
        edges = np.load(cache)
    else:
        edges = geometric_graph(n, deg, seed)[0] if family == "geometric" else heavy_tailed_graph(n, deg, seed=seed)[0]
        if world > 1:
            tmp = cache + f".tmp{os.getpid()}.npy"
            np.save(tmp, edges)
            os.replace(tmp, cache)          # atomic: other ranks either see the whole file or none
    wembed = host.load()
    t0 = time.perf_counter()
    graph = wembed.graphFromEdgeArray(edges)
    ingest_s = time.perf_counter() - t0
    rp, col = graph.csr()
    assert graph.getNumVertices() <= n
    if graph.getNumVertices() < n:         # trailing isolated vertices (n = largest id + 1 in the reference, Graph.cpp:101)
        rp = np.concatenate([rp, np.full(n - graph.getNumVertices(), rp[-1], np.int32)])
    w = degree_weights(n, edges, d)
    x0 = initial_coordinates(n, d, seed=1234)
    return dict(name=name, n=n, d=d, m=len(edges), edges=edges, row_ptr=rp, col=col, weights=w, x0=x0, ingest_s=ingest_s)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


class native_stdout_to_stderr:
    """The reference's C++ logs warnings with std::cout; stdout of this script carries exactly one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` captures summarised in profiles/r2_summary.md
# (c3; the window of the trajectory each capture was taken in is part of the record)
NCU_TRAFFIC = {
    "repel": {"bytes": 46.96e6 + 5.59e6, "window": "step 15 of c3 (profiles/r2_ncu_step15.csv); step 100: 103.9 + 20.7 MB"},
    "attract_update": {"bytes": 171.8e6 + 114.6e6, "window": "step 15 of c3 (profiles/r2_ncu_step15.csv); step 100: 180.2 + 117.9 MB"},
}


def algorithmic_bytes_per_step(n, m, d):
    """B_alg = 8m + 24n + 36nd (fp32 state, int32 ids, every array moved once)."""
    return 8 * m + 24 * n + 36 * n * d


def timed_async_steps(dev, first_it, steps):
    """K asynchronous steps, CUDA events on the handle's stream; returns (seconds, stats of every step)."""
    dev.mark(0)
    stats, inflight, it = [], 0, first_it
    for _ in range(steps):
        it += 1
        dev.step_async(lr_schedule(it))
        inflight += 1
        if inflight >= 32:
            stats.append(dev.step_collect())
            inflight -= 1
    dev.mark(1)
    while inflight:
        stats.append(dev.step_collect())
        inflight -= 1
    return dev.elapsed_ms(0, 1) * 1e-3, stats


def run_ours(args):
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # stdout carries exactly one JSON line
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    from wembed_b200 import build, cabi
    build.build()
    sampler = ClockSampler(local)
    sampler.start()                         # runs across all timed passes below
    wl = make_workload(args.workload, rank, world)
    n, d, m = wl["n"], wl["d"], wl["m"]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def fresh(w=None, **opts):
        """A new handle advanced by the W warm-up steps: every measurement below starts from the same layout AND the same
        optimizer state (the Adam moments cannot be restored through the ABI, so the warm-up is simply repeated)."""
        w = wl if w is None else w
        dev = cabi.DeviceEmbedder(w["row_ptr"], w["col"], embedding_dimension=w["d"], device=local, seed=1234, **opts)
        dev.set_weights(w["weights"])
        dev.set_coordinates(w["x0"])
        if world > 1:   # one graph, vertices range-partitioned over the GPUs, every rank stores its results into all replicas (NVLink)
            from wembed_b200 import sharding
            sharding.shard_embedder(dev, rank, world, torch.device("cuda", local))
        for i in range(1, args.warmup + 1):
            dev.step(lr_schedule(i))
        return dev

    def max_over_ranks(seconds):
        t = torch.tensor([seconds], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- e2e: blocking C ABI, host buffers in and out ---------------------------------------------------------
    dev = fresh()
    x_start = dev.coordinates()          # layout at the start of the timed window (also the CPU baseline's input)
    it = args.warmup
    barrier()
    t0 = time.perf_counter()
    dev.mark(2)
    dev.set_coordinates(x_start)
    t_set = time.perf_counter() - t0
    for _ in range(args.steps):
        it += 1
        dev.step(lr_schedule(it))
    t_steps = time.perf_counter() - t0 - t_set
    x_end = dev.coordinates()
    dev.mark(3)
    de_wall = time.perf_counter() - t0
    de_events = dev.elapsed_ms(2, 3) * 1e-3
    barrier()
    de = max_over_ranks(max(de_wall, de_events))   # host-visible time of the blocking calls (>= the device time)
    e2e_parts = {"set_coordinates_s": t_set, "steps_s": t_steps, "get_coordinates_s": de_wall - t_set - t_steps, "events_s": de_events}
    assert np.isfinite(x_end).all()
    dev.close()

    # ---- value: device-resident, K asynchronous steps, CUDA events on the handle's stream ---------------------
    dev = fresh()
    barrier()
    launches0 = dev.launch_count()
    dt, stats = timed_async_steps(dev, args.warmup, args.steps)
    launches = dev.launch_count() - launches0
    barrier()
    dt = max_over_ranks(dt)
    dev.close()

    # ---- per-phase device times (CUDA events around each phase) over the same window; the first step of this pass also gives the
    # device's half of the parity record (forces of the window's first step)
    dev = fresh(keep_forces=1)
    dev.enable_timing(True)
    it = args.warmup
    phases, f_first, st_first = [], None, None
    for k in range(args.steps):
        it += 1
        st = dev.step(lr_schedule(it))
        phases.append(dev.phase_times())
        if k == 0 and world == 1 and not args.no_cpu:
            f_first, st_first = dev.forces(), st
    ph = {k: float(np.mean([p[k] for p in phases])) for k in phases[0]}
    dev.close()

    ranks_identical = None
    if world > 1:   # every rank must hold the same statistics bit for bit (the pytest cases for this need a multi-GPU lease)
        key = [stats[-1][k] for k in ("loss_attract", "loss_repel", "rel_displacement", "num_repulsion_pairs", "sum_displacement")]
        mine = torch.tensor(key, device="cuda", dtype=torch.float64)
        allk = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allk, mine)
        ranks_identical = all(bool((a == allk[0]).all()) for a in allk)

    secondary = None
    if not args.no_secondary and args.workload == "c3":
        try:
            secondary = secondary_c5(args, fresh, barrier, max_over_ranks, rank, world)
        except Exception as e:   # never lose the headline line to the secondary one
            secondary = {"workload": "c5", "error": f"{type(e).__name__}: {e}"[:300]}

    clocks = sampler.stop()
    if rank != 0:
        return
    units = 2.0 * m * args.steps          # one graph in total, however many GPUs share it
    peak, peak_src = measured_peak_gbs()
    dom = max(("index", "attract_update", "repel", "recentre_observe"), key=lambda k: ph[k])
    bytes_step = algorithmic_bytes_per_step(n, m, d)
    V4 = 4 * ((d + 3) // 4)                     # padded row length
    kernel_bytes = {  # algorithmic bytes per launch of each kernel group
        # repulsion search: sorted points + ids + iw read once (the walk re-reads tree nodes from L1 / L2), found pairs written
        "repel": 4 * V4 * n + 8 * n,
        # SURVEY 8(d)'s count for the fused step kernel: CSR (rowPtr + col), iw, x / m / v read, x / m / v written
        "attract_update": 8 * m + 4 * n + 4 * n + 24 * V4 * n,
        # x read for the keys, key/value sort passes, x gathered into sorted planes, boxes written
        "index": 4 * V4 * n + 4 * 16 * n + 4 * V4 * n * 2 + 8 * n,
        "recentre_observe": 3 * 4 * V4 * n,
    }
    if world > 1:     # a rank's kernels move its share of the vertices (the index build is replicated on every rank); peak = one GPU's
        kernel_bytes = {k: (v if k == "index" else v // world) for k, v in kernel_bytes.items()}
        bytes_step = (bytes_step - kernel_bytes["index"]) // world + kernel_bytes["index"]
    rooflines = {k: {"algorithmic_bytes": kernel_bytes[k], "ms": ph[k], "achieved_gbs": kernel_bytes[k] / (ph[k] * 1e-3) / 1e9,
                     "frac": kernel_bytes[k] / (ph[k] * 1e-3) / 1e9 / peak} for k in kernel_bytes}
    dom_bytes = kernel_bytes[dom]
    achieved = dom_bytes / (ph[dom] * 1e-3) / 1e9
    window = f"steps {args.warmup + 1}..{args.warmup + args.steps}"
    out = {
        "metric": "edge_force_updates_per_s", "value": units / dt, "unit": "directed-edge force updates/s",
        "steps_per_s": args.steps / dt, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {WORKLOADS[args.workload][3]} graph n={n} m={m} d={d}, default options, "
                               f"trajectory {window} from the uniform-cube layout",
                   "parallelism": (f"{world} GPU(s): vertices range-partitioned, repulsion queries dealt by blocks of the sorted order; found pairs, "
                                   "per-block sums and recentred positions are stored straight into the consumers' memory over NVLink (CUDA IPC), "
                                   "three flag barriers per step, no collective on the data path") if world > 1 else "1 GPU",
                   "l2": "working set (x, m, v, CSR, index: ~230 MB at c3) exceeds the 126 MB L2; no flush needed"},
        "e2e": {"value": units / de, "unit": "directed-edge force updates/s", "steps_per_s": args.steps / de,
                "h2d_bytes_per_step": n * V4 * 4 / args.steps + 16, "d2h_bytes_per_step": n * V4 * 4 / args.steps + 8 * (14 + V4),
                "what": ("wb_set_coordinates(host doubles) + K blocking wb_step (observables copied to the host every step) + wb_get_coordinates(host "
                         "doubles); the coordinates cross the boundary ONCE per K steps (a device-resident loop, like calculateEmbedding), converted "
                         "to / from the device's fp32 rows on the host side of pinned staging buffers, so the per-step byte counts are n*d*4 / K"),
                "parts": e2e_parts},
        "gpu_launches": int(launches),
        # P * steps/s, P = repulsive pairs evaluated with a non-zero force per step (directed, counted by the fused kernel)
        "repulsive_pair_evals_per_s": float(np.mean([st.get("num_repulsion_pairs", 0.0) for st in stats])) * args.steps / dt,
        "phases_ms": ph,
        "roofline": {"bound": "hbm", "limited_by": ("instruction issue + the SM's L1 data path, NOT HBM: ncu of this kernel shows DRAM < 1 % of peak, L2 hit 99 %, L1 data-pipe "
                                                    "wavefronts ~85 %, issue slots ~77 % busy (profiles/); its HBM roofline fraction is therefore tiny by construction"),
                     "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_TRAFFIC.get(dom, {}).get("bytes"), "traffic_window": NCU_TRAFFIC.get(dom, {}).get("window"), "window": window,
                     "peak_source": peak_src, "algorithmic_bytes": dom_bytes, "per": "rank (one GPU's kernels against one GPU's peak)",
                     "kernels": rooflines, "fused_step_kernel": dict(rooflines["attract_update"], traffic=NCU_TRAFFIC["attract_update"]["bytes"],
                                                                     traffic_window=NCU_TRAFFIC["attract_update"]["window"]),
                     "whole_step": {"algorithmic_bytes": bytes_step, "achieved": bytes_step / (ph["total"] * 1e-3) / 1e9,
                                    "frac": bytes_step / (ph["total"] * 1e-3) / 1e9 / peak}},
        "clocks": clocks,
        "ingest": {"what": "wembed::graphFromEdges on the workload's edge array (sort-based CSR build of the product)", "edges": m, "seconds": wl["ingest_s"]},
        "last_step": {k: stats[-1][k] for k in ("loss_attract", "loss_repel", "rel_displacement", "num_repulsion_pairs", "num_candidates")},
    }
    if ranks_identical is not None:
        out["ranks_identical"] = ranks_identical
    if secondary is not None:
        out["secondary"] = secondary
    if world == 1 and not args.no_cpu:
        out["cpu_baseline"], out["parity_at_config"] = cpu_baseline(wl, x_start, args.warmup, f_first, st_first)
        out["same_sample"] = same_sample(args)
        out["convergence"] = convergence(wl)
    print(json.dumps(out), flush=True)


def secondary_c5(args, fresh, barrier, max_over_ranks, rank, world):
    """BASELINE.json configs[4] (n = 1e7, ~1e8 edges, d = 16) on the same GPUs: ms / step over a short window."""
    wl5 = make_workload("c5", rank, world)
    steps = 5
    dev = fresh(wl5)
    barrier()
    dt, stats = timed_async_steps(dev, args.warmup, steps)
    barrier()
    dt = max_over_ranks(dt)
    dev.close()
    return {"workload": f"c5: geometric graph n={wl5['n']} m={wl5['m']} d={wl5['d']}, steps {args.warmup + 1}..{args.warmup + steps}",
            "n_gpus": world, "ms_per_step": dt / steps * 1e3, "steps_per_s": steps / dt,
            "value": 2.0 * wl5["m"] * steps / dt, "unit": "directed-edge force updates/s",
            "last_step": {k: stats[-1][k] for k in ("loss_attract", "loss_repel", "num_repulsion_pairs")}}


def cpu_baseline(wl, x_start, iteration, f_dev, st_dev):
    """Three steps of the oracle port from the GPU's state at the start of the timed window (bounded sample); its first step is also the
    CPU half of the parity record: same layout, forces of the window's first step on both sides."""
    import oracle
    oracle.build("port")
    n, d, m = wl["n"], wl["d"], wl["m"]
    cores = os.cpu_count() or 1
    steps = 3
    with native_stdout_to_stderr():
        cpu = oracle.CpuEmbedder("port", wl["edges"], n=n, embeddingDimension=d, init_state=False, numThreads=cores)
        cpu.set_weights(wl["weights"])
        cpu.set_coordinates(x_start)
        cpu.step()                                   # un-timed: first touch of every buffer, and the parity sample
        f_cpu, cs = cpu.forces(), cpu.stats()
        t0 = time.perf_counter()
        for _ in range(steps):
            cpu.step()
        dt = time.perf_counter() - t0
        cpu.close()
    base = {"value": 2.0 * m * steps / dt, "unit": "directed-edge force updates/s", "steps_per_s": steps / dt, "cores": cores, "kind": "port",
            "sample": f"{steps} steps (after one un-timed step) of the same workload from the device state after {iteration} steps "
                      "(oracle/wembed_port.cpp, OpenMP, fp64)"}
    parity = None
    if f_dev is not None:
        scale = float(np.abs(f_cpu).max())
        err = np.abs(f_cpu - f_dev).max(axis=1) / scale
        parity = {"what": "first step of the timed window, device (fp32) vs oracle port (fp64), same layout",
                  "pairs_device": st_dev["num_repulsion_pairs"], "pairs_cpu": cs["num_rep_pairs"],
                  "pairs_equal": bool(st_dev["num_repulsion_pairs"] == cs["num_rep_pairs"]),
                  "loss_attract_rel_err": abs(st_dev["loss_attract"] - cs["loss_attract"]) / max(abs(cs["loss_attract"]), 1e-300),
                  "loss_repel_rel_err": abs(st_dev["loss_repel"] - cs["loss_repel"]) / max(abs(cs["loss_repel"]), 1e-300),
                  "max_force_rel_err": float(err.max()), "vertices_above_1e-5": int((err > 1e-5).sum()),
                  "note": "force error relative to the step's largest force component; a vertex above 1e-5 owns a pair within 1e-5 of the hinge "
                          "(|f| jumps by ws there, WembedEmbedder.cpp:163-168), tests/test_gpu_parity_at_size.py masks and counts them"}
    return base, parity


def same_sample(args):
    """This build on the SAMPLE the reference arm times (n = 20 000, same generator, d and window), so a same-configuration ratio can be
    formed from the two arms' lines."""
    from wembed_b200 import cabi
    n, deg, d, _ = WORKLOADS["small"]
    wl = make_workload("small")
    steps, warm = min(args.steps, 10), min(args.warmup, 5)
    dev = cabi.DeviceEmbedder(wl["row_ptr"], wl["col"], embedding_dimension=d, seed=1234)
    dev.set_weights(wl["weights"])
    dev.set_coordinates(wl["x0"])
    for i in range(1, warm + 1):
        dev.step(lr_schedule(i))
    dt, _ = timed_async_steps(dev, warm, steps)
    dev.close()
    return {"workload": f"geometric graph n={n} m={wl['m']} d={d} (the reference arm's bounded sample), steps {warm + 1}..{warm + steps}",
            "value": 2.0 * wl["m"] * steps / dt, "unit": "directed-edge force updates/s", "steps_per_s": steps / dt}


def convergence(wl):
    """What a user of wembed::Embedder sees: calculateEmbedding to convergence through the public C++ / Python API (default options,
    loss stop criterion, the embedder's own random layout), wall clock."""
    from wembed_b200 import host
    wembed = host.load()
    wembed.setSeed(1234)
    opts = wembed.Options()
    opts.embeddingDimension = wl["d"]
    g = wembed.graphFromEdgeArray(wl["edges"])
    t0 = time.perf_counter()
    emb = wembed.createEmbedder(g, opts)
    t_create = time.perf_counter() - t0
    iters = 0
    t0 = time.perf_counter()
    while not emb.isFinished():
        emb.calculateStep()
        iters += 1
    t_run = time.perf_counter() - t0
    loss = emb.getLoss()
    return {"what": "wembed::createEmbedder + calculateStep until isFinished (WembedEmbedder.cpp:65-86), default options",
            "iterations": iters, "seconds": t_run, "create_seconds": t_create, "final_loss": loss.total}


def run_reference(args):
    """The reference's own C++ (oracle/_ref, its SNN index) on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import oracle
    from wembed_b200.datasets import degree_weights, geometric_graph, heavy_tailed_graph, initial_coordinates
    n_full, deg, d, family = WORKLOADS[args.workload]
    kind = "reference" if oracle.have("ref") or oracle.build("ref") else "port"
    if kind == "port":
        oracle.build("port")
    n = min(n_full, WORKLOADS["small"][0])
    edges = geometric_graph(n, deg, 42)[0] if family == "geometric" else heavy_tailed_graph(n, deg, seed=42)[0]
    cores = os.cpu_count() or 1
    with native_stdout_to_stderr():
        cpu = oracle.CpuEmbedder("ref" if kind == "reference" else "port", edges, n=n, embeddingDimension=d, init_state=False, numThreads=cores)
        cpu.set_weights(degree_weights(n, edges, d))
        cpu.set_coordinates(initial_coordinates(n, d, seed=1234))
        for _ in range(args.warmup):
            cpu.step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu.step()
        dt = time.perf_counter() - t0
    m = len(edges)
    value = 2.0 * m * args.steps / dt
    sample = (f"{family} graph n={n} m={m} d={d} (same generator and options as the workload, smaller n: the reference's SNN index "
              f"scans O(n^(1-1/d)) points per query and cannot finish a step at n={n_full}); throughput is per directed edge")
    print(json.dumps({
        "impl": "reference", "metric": "edge_force_updates_per_s", "value": value, "unit": "directed-edge force updates/s",
        "steps_per_s_on_sample": args.steps / dt, "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload} (bounded sample)"},
        "cpu_baseline": {"value": value, "unit": "directed-edge force updates/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "directed-edge force updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity / same_sample / convergence legs")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary record: BASELINE.json's 100M-edge configuration (c5) on the same GPUs")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 10:
            args.steps = 10      # bounded: the whole reference run must end within minutes
        if args.warmup > 5:
            args.warmup = 5
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
