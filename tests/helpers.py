"""Shared helpers of the parity tests."""
from __future__ import annotations

import numpy as np

from wembed_b200.datasets import degree_weights, geometric_graph, initial_coordinates


def ring_graph(n=64):
    """The graph of the reference's tests/TestDeterminism.cpp:15-22: a ring with +1 and +7 chords."""
    e = []
    for v in range(n):
        e.append((v, (v + 1) % n))
        e.append((v, (v + 7) % n))
    return np.asarray(e, np.int32)


SMALL_GRAPH = np.asarray([(0, 1), (1, 2), (2, 3), (3, 4), (1, 3), (2, 4)], np.int32)  # assets/small_graph.edg


def lr_exponential(iteration, lr0=10.0, cooling=0.995, warmup=20):
    """LRScheduler::learningRate with ExponentialCoolingSchedule (LRScheduler.cpp:7-17)."""
    lr = lr0 * cooling ** float(iteration)
    return lr * iteration / warmup if iteration < warmup else lr


def near_threshold_vertices(x, w, row_ptr, col, L=1.0, tau=1e-5, max_pairs=50_000_000):
    """Vertices owning a pair whose weighted distance dist*ws lies within tau*L of the hinge at L.

    fp32 and fp64 may legitimately disagree on which side of the discontinuity such a pair falls
    ("hinge discontinuity"); parity tests mask these vertices and report their count.
    """
    from scipy.spatial import cKDTree
    n, d = x.shape
    iw = w ** (-1.0 / d)
    src = np.repeat(np.arange(n), np.diff(row_ptr))
    dist = np.sqrt(((x[src] - x[col]) ** 2).sum(1))
    near = np.abs(dist * iw[src] * iw[col] - L) <= tau * L
    flagged = np.zeros(n, bool)
    flagged[src[near]] = True
    rmax = L * (w.max() ** 2) ** (1.0 / d) * (1 + 2 * tau)
    tree = cKDTree(x)
    # only a thin shell can be near the threshold, but the shell radius depends on the pair: take all pairs
    # within rmax (bounded) and test them exactly
    pairs = tree.query_pairs(rmax, output_type="ndarray")
    if len(pairs) > max_pairs:
        raise RuntimeError("state too dense for the near-threshold scan")
    if len(pairs):
        a, b = pairs[:, 0], pairs[:, 1]
        dist = np.sqrt(((x[a] - x[b]) ** 2).sum(1))
        near = np.abs(dist * iw[a] * iw[b] - L) <= tau * L
        flagged[a[near]] = True
        flagged[b[near]] = True
    return flagged


def make_problem(n, d, avg_degree=10.0, seed=42):
    edges, _ = geometric_graph(n, avg_degree, seed)
    w = degree_weights(n, edges, d)
    x0 = initial_coordinates(n, d, seed=1234)
    return edges, w, x0


# ---- test-side restatement of the host scalar logic (used until a test drives the C++ facade instead) ----------
class LossMonitor:
    """ConvergenceMonitor (src/embeddingLib/src/embedder/ConvergenceMonitor.cpp:6-42)."""

    def __init__(self, tol=1e-3, patience=50, alpha=0.3, window=30):
        self.tol, self.patience, self.alpha = tol, patience, alpha
        self.ring = [0.0] * (max(1, window) + 1)
        self.head = self.count = self.observed = self.stagnant = 0
        self.smoothed, self.rate = 0.0, float("inf")

    def observe(self, loss):
        self.smoothed = loss if self.observed == 0 else self.alpha * loss + (1 - self.alpha) * self.smoothed
        self.observed += 1
        self.ring[self.head] = self.smoothed
        self.head = (self.head + 1) % len(self.ring)
        self.count = min(self.count + 1, len(self.ring))
        if self.count >= len(self.ring):
            start = self.ring[self.head]
            self.rate = (start - self.smoothed) / max(abs(start), 1e-12)
        else:
            self.rate = float("inf")
        self.stagnant = self.stagnant + 1 if self.rate < self.tol else 0

    def converged(self):
        return self.stagnant >= self.patience


def run_to_convergence(dev, o):
    """WembedEmbedder::calculateEmbedding loop (WembedEmbedder.cpp:65-86) over a DeviceEmbedder; ExponentialCooling only."""
    mon = LossMonitor(o.get("stopLossTol", 1e-3), o.get("stopLossPatience", 50), o.get("lossSmoothingFactor", 0.3), o.get("lossRateWindow", 30))
    settled, it, st = 0, 0, None
    while it < o.get("maxIterations", 10000):
        if o.get("stopCriterion", 1) == 0 and settled >= o.get("stopDisplacementPatience", 5):
            break
        if o.get("stopCriterion", 1) == 1 and mon.converged():
            break
        it += 1
        st = dev.step(lr_exponential(it, o.get("learningRate", 10.0), o.get("lrCoolingFactor", 0.995), o.get("warmupSteps", 20)))
        settled = settled + 1 if st["rel_displacement"] < o.get("stopDisplacementTol", 3e-4) else 0
        mon.observe(st["loss_attract"] + st["loss_repel"])
    return it, st


def reconstruction_metrics(x, w, row_ptr, col, nodes=None):
    """constructDeg (precision at k = deg) and MAP of the reference's evaluationLib on the WeightedGeometric similarity
    dist / (w_a w_b)^(1/d)  (src/evaluationLib/src/metrics/NodeSampler.cpp:5-111, Reconstruction.cpp:6-23,
    src/embeddingLib/src/embeddingSpace/WeightedGeometric.cpp:17-21).  Ties are broken by node id like the reference's
    sort of (similarity, id) pairs.  `nodes` = the sampled vertices (all by default); isolated vertices are skipped."""
    n, d = x.shape
    iw = w ** (-1.0 / d)
    nodes = np.arange(n) if nodes is None else np.asarray(nodes)
    deg_prec, avg_prec = [], []
    for v in nodes:
        nb = col[row_ptr[v]:row_ptr[v + 1]]
        if len(nb) == 0:
            continue
        sim = np.sqrt(((x - x[v]) ** 2).sum(1)) * iw * iw[v]
        order = np.lexsort((np.arange(n), sim))
        order = order[order != v]
        is_nb = np.zeros(n, bool)
        is_nb[nb] = True
        hits = is_nb[order]
        prec = np.cumsum(hits) / np.arange(1, n)
        deg_prec.append(prec[len(nb) - 1])
        avg_prec.append(prec[hits].mean())
    return float(np.mean(deg_prec)), float(np.mean(avg_prec))


def edge_detection_metrics(x, w, v, u, is_edge, n, m):
    """EdgeDetection::getMetricValues (src/evaluationLib/src/metrics/EdgeDetection.cpp:6-66) on the WeightedGeometric
    similarity, restated loop for loop (running percentages, strict `F1 > best`); the pairs are sorted by similarity like
    EdgeSampler.cpp:62 (stable here, so ties keep the sampler's order).  Returns (precision, recall, F1)."""
    d = x.shape[1]
    iw = w ** (-1.0 / d)
    v, u, is_edge = np.asarray(v), np.asarray(u), np.asarray(is_edge).astype(bool)
    wr = w ** (1.0 / d)
    sim = np.sqrt(((x[v] - x[u]) ** 2).sum(1)) / (wr[v] * wr[u])          # WeightedGeometric.cpp:17-21
    order = np.argsort(sim, kind="stable")
    n_e, n_ne = int(is_edge.sum()), int((~is_edge).sum())
    M, no_m = float(m), float(n * (n - 1) // 2 - m)
    wrong_e, wrong_ne = 1.0, 0.0
    best = (-1.0, -1.0, -1.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in order:
            if is_edge[i]:
                wrong_e -= 1.0 / n_e
            else:
                wrong_ne += 1.0 / n_ne
            tp = (1.0 - wrong_e) * M
            retrieved = tp + wrong_ne * no_m
            precision, recall = np.float64(tp) / np.float64(retrieved), np.float64(tp) / np.float64(M)
            f1 = np.float64(2.0) / (np.float64(1.0) / precision + np.float64(1.0) / recall)
            if f1 > best[2]:
                best = (float(precision), float(recall), float(f1))
    return best


def brute_force_repulsive_pairs(x, w, row_ptr, col, L=1.0, tau=1e-5, tile=2048, device="cuda"):
    """Index-free census of the repulsive pairs of a layout: ALL n^2 ordered pairs are tested in tiles on the GPU with plain
    torch (a matmul-form distance in fp32 selects a generous superset, which is then re-evaluated in float64), so nothing of
    the product's index, tree or walk is shared with it.

    A pair (v, u), u not in N(v), u != v, counts when dist * ws <= L, ws = (w_v w_u)^(-1/d) (WembedEmbedder.cpp:196-201).
    Returns (lo, hi, degree_lo, degree_hi): directed pair counts with dist * ws <= L (1 - tau) / <= L (1 + tau) - an
    implementation that evaluates the predicate in fp32 must land between the two - and the same per vertex.
    """
    import torch
    n, d = x.shape
    dev = torch.device(device)
    x64 = torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device=dev)
    xc = (x64 - x64.mean(0, keepdim=True)).to(torch.float32)
    iw64 = torch.as_tensor(w ** (-1.0 / d), dtype=torch.float64, device=dev)
    iwsq = (iw64 * iw64).to(torch.float32)
    nb = (xc * xc).sum(1)
    src = np.repeat(np.arange(n, dtype=np.int64), np.diff(row_ptr))
    ekeys = torch.as_tensor(src * n + col.astype(np.int64), device=dev)       # CSR order = sorted by (src, dst)
    lo_deg = torch.zeros(n, dtype=torch.int64, device=dev)
    hi_deg = torch.zeros(n, dtype=torch.int64, device=dev)
    thr = float(L * L * 1.02)
    for a0 in range(0, n, tile):
        a1 = min(n, a0 + tile)
        g = xc[a0:a1] @ xc.T
        g.mul_(-2.0).add_(nb[a0:a1, None]).add_(nb[None, :])
        g.mul_(iwsq[a0:a1, None]).mul_(iwsq[None, :])
        ri, cj = (g <= thr).nonzero(as_tuple=True)
        del g
        ri = ri + a0
        keep = ri != cj
        ri, cj = ri[keep], cj[keep]
        if ri.numel() == 0:
            continue
        val = (x64[ri] - x64[cj]).pow(2).sum(1).sqrt() * iw64[ri] * iw64[cj]
        key = ri * n + cj
        pos = torch.searchsorted(ekeys, key).clamp_(max=max(ekeys.numel() - 1, 0))
        is_nb = (ekeys[pos] == key) if ekeys.numel() else torch.zeros_like(key, dtype=torch.bool)
        ok_lo = (~is_nb) & (val <= L * (1.0 - tau))
        ok_hi = (~is_nb) & (val <= L * (1.0 + tau))
        lo_deg += torch.bincount(ri[ok_lo], minlength=n)
        hi_deg += torch.bincount(ri[ok_hi], minlength=n)
    return int(lo_deg.sum()), int(hi_deg.sum()), lo_deg.cpu().numpy(), hi_deg.cpu().numpy()


def near_threshold_edge_owners(x, w, row_ptr, col, L=1.0, tau=1e-5, chunk=1 << 21):
    """Vertices owning a graph edge whose weighted length lies within tau * L of the attraction hinge (chunked numpy)."""
    n, d = x.shape
    iw = w ** (-1.0 / d)
    src = np.repeat(np.arange(n, dtype=np.int64), np.diff(row_ptr))
    flagged = np.zeros(n, bool)
    for a in range(0, len(col), chunk):
        s, c = src[a:a + chunk], col[a:a + chunk]
        dist = np.sqrt(((x[s] - x[c]) ** 2).sum(1))
        flagged[s[np.abs(dist * iw[s] * iw[c] - L) <= tau * L]] = True
    return flagged
