"""Synthetic benchmark inputs (host side, numpy only).

Two graph families, following SURVEY.md section 8(d):

* geometric: the family of the reference's GeometricGraphSampler
  (src/graphLib/src/graph/GeometricGraphSampler.cpp:10-51): n points uniform in
  [0, sqrt(n)]^2, an edge iff the Euclidean distance is below sqrt(avg_degree / pi).
  All vertices are kept (no largest-component filter) so n is exact.
* heavy_tailed: a threshold GIRG in the spirit of src/cli_generator/GirgGenerator.cpp:14-44
  (the external `girgs` library is not available offline): power-law weights
  w_i = u_i^(-1/(beta-1)), positions uniform on the unit torus-free square, an edge iff
  dist < c * sqrt(w_i w_j / W); c is tuned by bisection to the requested average degree.

Both are built with a uniform grid (O(n) expected), not the reference's O(n^2) loop.
Returned edges are undirected, each once, as an int32 array of shape (m, 2) with src < dst.
"""
from __future__ import annotations

import subprocess

import numpy as np


def _pairs_within(points: np.ndarray, radius: float) -> np.ndarray:
    """All index pairs (i < j) with ||p_i - p_j|| < radius, via a cell grid of side `radius`."""
    n = len(points)
    if n == 0:
        return np.zeros((0, 2), np.int32)
    cell = np.floor(points / radius).astype(np.int64)
    cell -= cell.min(axis=0)
    ny = int(cell[:, 1].max()) + 3
    key = (cell[:, 0] + 1) * ny + (cell[:, 1] + 1)
    order = np.argsort(key, kind="stable")
    skey = key[order]
    ncell = int((cell[:, 0].max() + 3) * ny)
    start = np.searchsorted(skey, np.arange(ncell + 1))
    out = []
    r2 = radius * radius
    # half stencil so every unordered cell pair is visited once
    for dx, dy in ((0, 0), (0, 1), (1, -1), (1, 0), (1, 1)):
        nkey = skey + dx * ny + dy
        lo, hi = start[nkey], start[nkey + 1]
        cnt = hi - lo
        if dx == 0 and dy == 0:
            pass
        tot = int(cnt.sum())
        if tot == 0:
            continue
        a = np.repeat(np.arange(n), cnt)
        offs = np.arange(tot) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        b = np.repeat(lo, cnt) + offs
        if dx == 0 and dy == 0:
            keep = b > a
            a, b = a[keep], b[keep]
        ia, ib = order[a], order[b]
        dv = points[ia] - points[ib]
        ok = (dv * dv).sum(axis=1) < r2
        out.append(np.stack([ia[ok], ib[ok]], axis=1))
    e = np.concatenate(out) if out else np.zeros((0, 2), np.int64)
    e = np.sort(e, axis=1)
    e = e[np.lexsort((e[:, 1], e[:, 0]))]
    return e.astype(np.int32)


def geometric_graph(n: int, avg_degree: float = 10.0, seed: int = 42):
    """2-D random geometric graph; returns (edges[m,2] int32, points[n,2] float64)."""
    rng = np.random.default_rng(seed)
    pts = rng.random((n, 2)) * np.sqrt(n)
    radius = float(np.sqrt(avg_degree / np.pi))
    try:                                   # C++ helper: same pairs in the same order, ~50x faster at n = 1e7
        from . import datagen
        return datagen.pairs_within(pts, radius), pts
    except (OSError, subprocess.CalledProcessError):
        return _pairs_within(pts, radius), pts


def _pairs_between(P: np.ndarray, Q: np.ndarray, radius: float, same: bool) -> np.ndarray:
    """Index pairs (i in P, j in Q) with distance < radius (i < j only when P is Q), grid on Q."""
    if len(P) == 0 or len(Q) == 0:
        return np.zeros((0, 2), np.int64)
    lo = np.minimum(P.min(axis=0), Q.min(axis=0))
    cq = np.floor((Q - lo) / radius).astype(np.int64) + 1
    cp = np.floor((P - lo) / radius).astype(np.int64) + 1
    ny = int(max(cq[:, 1].max(), cp[:, 1].max())) + 2
    nx = int(max(cq[:, 0].max(), cp[:, 0].max())) + 2
    qkey = cq[:, 0] * ny + cq[:, 1]
    order = np.argsort(qkey, kind="stable")
    start = np.searchsorted(qkey[order], np.arange(nx * ny + 1))
    pkey = cp[:, 0] * ny + cp[:, 1]
    out = []
    r2 = radius * radius
    chunk = 1 << 18
    for s0 in range(0, len(P), chunk):
        pk = pkey[s0:s0 + chunk]
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                nk = pk + dx * ny + dy
                lo_i, hi_i = start[nk], start[nk + 1]
                cnt = hi_i - lo_i
                tot = int(cnt.sum())
                if tot == 0:
                    continue
                a = np.repeat(np.arange(s0, s0 + len(pk)), cnt)
                b = order[np.repeat(lo_i, cnt) + (np.arange(tot) - np.repeat(np.cumsum(cnt) - cnt, cnt))]
                if same:
                    keep = b > a
                    a, b = a[keep], b[keep]
                dv = P[a] - Q[b]
                ok = (dv * dv).sum(axis=1) < r2
                out.append(np.stack([a[ok], b[ok]], axis=1))
    return np.concatenate(out) if out else np.zeros((0, 2), np.int64)


def heavy_tailed_graph(n: int, avg_degree: float = 20.0, beta: float = 2.5, seed: int = 42, c: float | None = None, native: bool = True):
    """Threshold GIRG-like graph with power-law weights; returns (edges, generator weights).

    The expected degree, pi * c^2 * E[w] up to boundary effects, does not depend on n, so for large n the constant c
    is calibrated once on a 100k-vertex instance of the same distribution."""
    if c is None and n > 200_000:
        c = _calibrate_c(avg_degree, beta, seed, native)
    rng = np.random.default_rng(seed)
    w = (1.0 - rng.random(n)) ** (-1.0 / (beta - 1.0))
    pts = rng.random((n, 2))
    W = w.sum()
    cls = np.floor(np.log2(w)).astype(np.int64)          # weight layers [2^k, 2^(k+1))
    layers = [np.flatnonzero(cls == k) for k in range(int(cls.max()) + 1)]

    girg = None
    if native:
        try:                               # C++ helper: the same edges in the same order, ~40x faster at n = 1e6
            from . import datagen
            girg = datagen.girg_pairs
        except (OSError, subprocess.CalledProcessError):
            girg = None

    def build(c, count_only):
        if girg is not None:
            return girg(pts, w, c, W, count_only)
        parts, total = [], 0
        for ka, A in enumerate(layers):
            for kb in range(ka, len(layers)):
                B = layers[kb]
                if len(A) == 0 or len(B) == 0:
                    continue
                rmax = min(1.5, c * np.sqrt(2.0 ** (ka + 1) * 2.0 ** (kb + 1) / W))
                cand = _pairs_between(pts[A], pts[B], float(rmax), ka == kb)
                ia, ib = A[cand[:, 0]], B[cand[:, 1]]
                dv = pts[ia] - pts[ib]
                ok = (dv * dv).sum(1) < (c * c) * w[ia] * w[ib] / W
                total += int(ok.sum())
                if not count_only:
                    parts.append(np.stack([ia[ok], ib[ok]], 1))
        if count_only:
            return total
        e = np.concatenate(parts) if parts else np.zeros((0, 2), np.int64)
        e = np.sort(e, axis=1)
        e = e[np.lexsort((e[:, 1], e[:, 0]))]
        return e.astype(np.int32)

    if c is not None:
        return build(c, False), w
    lo, hi = 0.0, 4.0 * np.sqrt(avg_degree)
    target = avg_degree * n / 2.0
    for _ in range(14):
        mid = 0.5 * (lo + hi)
        if build(mid, True) < target:
            lo = mid
        else:
            hi = mid
    heavy_tailed_graph.last_c = 0.5 * (lo + hi)
    return build(0.5 * (lo + hi), False), w


def _calibrate_c(avg_degree, beta, seed, native=True):
    heavy_tailed_graph(100_000, avg_degree, beta, seed, native=native)
    return heavy_tailed_graph.last_c


def degree_weights(n: int, edges: np.ndarray, d: int, dimension_hint: float = -1.0) -> np.ndarray:
    """WembedEmbedder::constructDegreeWeights + rescaleWeights (WembedEmbedder.cpp:359-390)."""
    deg = np.bincount(np.asarray(edges).ravel(), minlength=n).astype(np.float64)
    w = np.maximum(deg, 1.0)
    if dimension_hint > 0:
        w = w ** (float(d) / dimension_hint)
    s = 0.0
    for chunk in np.array_split(w, max(1, n // 65536)):  # sequential-sum order does not matter for the callers
        s += float(chunk.sum())
    return w * (float(n) / float(w.sum()))


def initial_coordinates(n: int, d: int, seed: int = 1234, fp32_exact: bool = True) -> np.ndarray:
    """Uniform in [0, n^(1/d))^d like EmbedderInterface::constructRandomCoordinates
    (EmbedderInterface.hpp:61-65).  With fp32_exact the values are rounded to float32 so an
    fp32 device state and the fp64 oracle start from identical numbers."""
    rng = np.random.default_rng(seed)
    x = rng.random((n, d)) * (float(np.float32(n)) ** (1.0 / d))
    if fp32_exact:
        x = x.astype(np.float32).astype(np.float64)
    return x
