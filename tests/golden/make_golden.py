"""Generates tests/golden/*.npz from the reference's own sources (oracle/_ref, built in place from /root/reference
by oracle/Makefile).  Run here, where the reference checkout exists:  python tests/golden/make_golden.py

The fixtures pin the CPU port (tests -m "not gpu") and, on the GPU box where /root/reference does not exist, the
device path (tests -m gpu).  Index = SNN (the only index whose source is in the reference tree).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from helpers import SMALL_GRAPH, ring_graph  # noqa: E402
from wembed_b200.datasets import geometric_graph  # noqa: E402


def trace(cpu, steps, forces=True):
    out = {"x0": cpu.coordinates(), "w": cpu.weights()}
    xs, fs, st = [], [], []
    for _ in range(steps):
        cpu.step()
        s = cpu.stats()
        xs.append(cpu.coordinates())
        if forces:
            fs.append(cpu.forces())
        st.append([s["loss_attract"], s["loss_repel"], s["lr"], s["rel_displacement"], s["rel_loss_improvement"], s["iteration"]])
    out["x"] = np.stack(xs)
    if forces:
        out["f"] = np.stack(fs)
    out["stats"] = np.asarray(st)
    return out


def main():
    assert oracle.build("ref"), "reference checkout not available"
    # (a) tests/TestDeterminism.cpp protocol: ring-64 (+1, +7 chords), d = 2, seed 1234, all-coincident start, 25 steps
    ring = ring_graph(64)
    cpu = oracle.CpuEmbedder("ref", ring, seed=1234, embeddingDimension=2, maxIterations=1000)
    cpu.set_coordinates(np.zeros((64, 2)))
    g = trace(cpu, 25)
    np.savez_compressed(os.path.join(HERE, "ring64_d2_coincident.npz"), edges=ring, **g)
    # (a') same graph, the constructor's random layout (Rand::setSeed(1234)), 10 steps
    cpu = oracle.CpuEmbedder("ref", ring, seed=1234, embeddingDimension=2, maxIterations=1000)
    g = trace(cpu, 10)
    np.savez_compressed(os.path.join(HERE, "ring64_d2_random.npz"), edges=ring, **g)
    # (b) assets/small_graph.edg, defaults (d = 4), to convergence
    for seed in (1, 2):
        cpu = oracle.CpuEmbedder("ref", SMALL_GRAPH, seed=seed)
        x0, w = cpu.coordinates(), cpu.weights()
        iters = cpu.run()
        s = cpu.stats()
        np.savez_compressed(os.path.join(HERE, f"small_graph_seed{seed}.npz"), edges=SMALL_GRAPH, x0=x0, w=w, iterations=iters,
                            x_final=cpu.coordinates(), loss_final=s["loss_attract"] + s["loss_repel"], csr_row=cpu.csr()[0], csr_col=cpu.csr()[1])
    # (c) geometric graph n = 600, d = 4 and d = 8: 6 steps with forces; candidate sets at the state after step 3
    for d in (4, 8):
        n = 600
        edges, _ = geometric_graph(n, 10, seed=5)
        cpu = oracle.CpuEmbedder("ref", edges, n=n, seed=99, embeddingDimension=d)
        x0 = cpu.coordinates().astype(np.float32).astype(np.float64)   # fp32-exact start
        cpu.set_coordinates(x0)
        g = trace(cpu, 3)
        queries = np.arange(0, n, 30, dtype=np.int32)
        cands = [np.sort(cpu.candidates(int(q))) for q in queries]
        g2 = trace(cpu, 3)
        rp, col = cpu.csr()
        np.savez_compressed(os.path.join(HERE, f"geo600_d{d}.npz"), edges=edges, x0=x0, w=g["w"], x=np.concatenate([g["x"], g2["x"]]),
                            f=np.concatenate([g["f"], g2["f"]]), stats=np.concatenate([g["stats"], g2["stats"]]), queries=queries,
                            cand_offsets=np.cumsum([0] + [len(c) for c in cands]), cand_ids=np.concatenate(cands), csr_row=rp, csr_col=col)
    # (d) option coverage on a small geometric graph: Simple optimizer, centre force, unit weights, dimension hint, scales
    n = 300
    edges, _ = geometric_graph(n, 8, seed=11)
    variants = {
        "simple": dict(optimizerType=0, simpleOptMaxDisplacement=0.5),
        "centre": dict(centreScale=0.05),
        "unit": dict(weightType=0),
        "hint": dict(dimensionHint=2.0, embeddingDimension=3),
        "scales": dict(attractionScale=2.0, repulsionScale=0.5, edgeLength=1.5),
        "adaptive": dict(lrScheduleType=1, lossRateWindow=3, lrAdaptPatience=2),
    }
    out = {}
    for name, o in variants.items():
        cpu = oracle.CpuEmbedder("ref", edges, n=n, seed=7, **o)
        x0 = cpu.coordinates().astype(np.float32).astype(np.float64)
        cpu.set_coordinates(x0)
        g = trace(cpu, 4)
        for k, v in g.items():
            out[f"{name}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "geo300_options.npz"), edges=edges, **out)
    # (e) multilevel driver: parent pointers of LabelPropagation::coarsenAllLayers and the outcome of a whole
    #     LayeredEmbedder run, straight from the reference through oracle/ref_harness.cpp
    import ctypes as C
    from oracle.oracle import OrcOptions, _PATHS
    from helpers import reconstruction_metrics
    from wembed_b200 import cabi
    from wembed_b200.datasets import heavy_tailed_graph
    ref = C.CDLL(_PATHS["ref"])
    ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    out = {}
    for name, e in (("ring64", ring), ("geo3000", geometric_graph(3000, 10, seed=5)[0]), ("heavy4000", heavy_tailed_graph(4000, 20, seed=1)[0])):
        src, dst = np.ascontiguousarray(e[:, 0], dtype=np.int32), np.ascontiguousarray(e[:, 1], dtype=np.int32)
        sizes, parents = np.zeros(64, np.int32), np.zeros(4 * len(e) + 64, np.int32)
        ref.ref_coarsen.restype = C.c_int32
        nl = ref.ref_coarsen(C.c_int64(len(src)), src.ctypes.data_as(ip), dst.ctypes.data_as(ip), sizes.ctypes.data_as(ip), 64,
                             parents.ctypes.data_as(ip), C.c_int64(len(parents)))
        out[f"{name}_edges"], out[f"{name}_sizes"], out[f"{name}_parents"] = e, sizes[:nl].copy(), parents[: sizes[:nl].sum()].copy()
        if name == "geo3000":
            n = int(e.max()) + 1
            o = OrcOptions()
            ref.ref_options_default(C.byref(o))
            o.embeddingDimension = 4
            x, w, st = np.zeros((n, 4)), np.zeros(n), np.zeros(8)
            ref.ref_layered_run.restype = C.c_int64
            iters = ref.ref_layered_run(C.c_int64(len(src)), src.ctypes.data_as(ip), dst.ctypes.data_as(ip), C.byref(o), 7, x.ctypes.data_as(dp),
                                        w.ctypes.data_as(dp), st.ctypes.data_as(dp))
            rp, col = cabi.csr_from_edges(n, e)
            q = reconstruction_metrics(x, w, rp, col, np.arange(0, n, 6))
            out["geo3000_layered"] = np.array([iters, st[0] + st[1], q[0], q[1]])
    np.savez_compressed(os.path.join(HERE, "hierarchy.npz"), **out)
    # (f) host scalar logic: the reference's own learning rates, loss rates and stop iterations on ring-64 (tests/TestDeterminism.cpp
    #     option sets) together with the loss / displacement sequences that produced them
    sched = {}
    cases = {
        "loss_stop": dict(maxIterations=2000, stopCriterion=1, lossRateWindow=10, stopLossTol=1e-2, stopLossPatience=10),
        "disp_stop": dict(maxIterations=5000, stopCriterion=0, stopDisplacementTol=1e-3, stopDisplacementPatience=5),
        "adaptive": dict(maxIterations=400, lrScheduleType=1, lossRateWindow=10, lrDecayThreshold=1e-2, lrDecayFactor=0.5,
                         lrGrowthThreshold=1e-1, lrGrowthFactor=1.0, lrAdaptPatience=5),
        "default": dict(maxIterations=300),
    }
    for name, o in cases.items():
        cpu = oracle.CpuEmbedder("ref", ring, seed=1234, embeddingDimension=2, **o)
        rows = []
        while not cpu.is_finished():
            cpu.step()
            s_ = cpu.stats()
            rows.append([s_["loss_attract"] + s_["loss_repel"], s_["rel_displacement"], s_["lr"], s_["rel_loss_improvement"]])
        op = cpu.opts
        sched[name + "_trace"] = np.asarray(rows)
        sched[name + "_opts"] = np.array([op.lrScheduleType, op.learningRate, op.warmupSteps, op.lrCoolingFactor, op.lrDecayFactor, op.lrDecayThreshold,
                                          op.lrAdaptPatience, op.lrGrowthFactor, op.lrGrowthThreshold, op.stopCriterion, op.stopDisplacementTol,
                                          op.stopDisplacementPatience, op.lossSmoothingFactor, op.lossRateWindow, op.stopLossTol, op.stopLossPatience,
                                          op.maxIterations], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "host_logic.npz"), **sched)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
