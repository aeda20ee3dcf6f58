"""Host-side plumbing of the vertex-sharded multi-GPU step (one process per GPU, torch.distributed for rendezvous).

The data path itself (owned-range kernels, NCCL all-gathers) lives in libwembed_b200.so; this module only
 * mirrors the library's vertex partition so callers can reason about ownership, and
 * ships the NCCL unique id from rank 0 to the other ranks over whatever torch.distributed backend is up
   (gloo on CPU in the tests, nccl on the GPU box).
"""
from __future__ import annotations

import numpy as np


def partition(n: int, world: int) -> list[tuple[int, int]]:
    """[begin, end) of every rank: contiguous ranges of ceil(n / world) vertices (wb_comm_init)."""
    rows = -(-max(n, 1) // world)
    return [(min(n, r * rows), min(n, r * rows + rows)) for r in range(world)]


def owner_of(v, n: int, world: int):
    rows = -(-max(n, 1) // world)
    return np.asarray(v) // rows


def exchange_unique_id(make_id, rank: int, world: int, device=None) -> bytes:
    """rank 0 calls make_id() (-> 128 bytes); every rank returns the same bytes."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return make_id()
    buf = torch.zeros(128, dtype=torch.uint8, device=device or "cpu")
    if rank == 0:
        raw = make_id()
        assert len(raw) == 128
        buf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().tolist())


def shard_embedder(dev, rank: int, world: int, device=None):
    """Joins `dev` (a cabi.DeviceEmbedder holding the full problem) to the sharded step."""
    from . import cabi
    uid = exchange_unique_id(cabi.comm_unique_id, rank, world, device)
    dev.comm_init(uid, rank, world)
    return partition(dev.n, world)[rank]
