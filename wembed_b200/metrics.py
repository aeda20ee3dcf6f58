"""Host-side sampling for the on-device quality metrics (mirrors of evaluationLib's samplers, numpy only).

The device evaluates the metrics (`wb_reconstruction`, `wb_edge_detection`); which vertices / pairs are scored is the
caller's choice, as it is in the reference, where the samplers draw from the global generator.
"""
from __future__ import annotations

import numpy as np


def sample_edge_pairs(row_ptr: np.ndarray, col: np.ndarray, scale: float = 10.0, seed: int = 0):
    """EdgeSampler::sampleHistEntries (src/evaluationLib/src/metrics/EdgeSampler.cpp:7-66) without the similarities.

    Returns (v, w, is_edge): every edge once (v < w, in CSR order, :22-30), then the non-edges the reference's walk over the
    n x n index space keeps: positions are visited with jumps of `geometric(p) + 1`, p = min(1, scale * m / (n(n-1)/2 - m)),
    and a position (v, w) is kept iff w > v and {v, w} is not an edge (:33-58).  The generator is numpy's, not the
    reference's std::mt19937 stream: the sampled sets agree in distribution, not element by element.
    """
    row_ptr = np.asarray(row_ptr, np.int64)
    col = np.asarray(col, np.int64)
    n = len(row_ptr) - 1
    src = np.repeat(np.arange(n, dtype=np.int64), np.diff(row_ptr))
    up = col > src
    ev, ew = src[up], col[up]
    m = len(ev)
    no_m = n * (n - 1) // 2 - m
    if n < 2 or no_m <= 0 or m == 0:
        return ev.astype(np.int32), ew.astype(np.int32), np.ones(m, np.uint8)
    p = min(1.0, scale * m / no_m)
    rng = np.random.default_rng(seed)
    total = n * n
    pos_parts, last = [], 0
    while last < total:
        # numpy's geometric counts trials (>= 1) = the reference's failures + 1
        jumps = rng.geometric(p, size=max(1024, int(1.1 * p * (total - last)) + 1024)).astype(np.int64)
        pos = last + np.cumsum(jumps)
        pos_parts.append(pos[pos < total])
        last = int(pos[-1])
    pos = np.concatenate(pos_parts)
    v, w = pos // n, pos % n
    keep = w > v
    v, w = v[keep], w[keep]
    # {v, w} an edge?  binary search of w in v's sorted CSR row, vectorised through global keys
    key = src * n + col
    q = v * n + w
    at = np.searchsorted(key, q)
    is_nb = (at < len(key)) & (key[np.minimum(at, len(key) - 1)] == q)
    v, w = v[~is_nb], w[~is_nb]
    return (np.concatenate([ev, v]).astype(np.int32), np.concatenate([ew, w]).astype(np.int32),
            np.concatenate([np.ones(m, np.uint8), np.zeros(len(v), np.uint8)]))
