"""Host-side plumbing of the vertex-sharded multi-GPU step (one process per GPU, torch.distributed for rendezvous).

The data path itself (owned-range kernels that store into the peers' memory over CUDA IPC mappings, flag barriers) lives in
libwembed_b200.so; this module only
 * mirrors the library's vertex partition so callers can reason about ownership (small graphs on many GPUs leave the last ranks
   without vertices: begin = end = n), and
 * ships the NCCL unique id from rank 0 to the other ranks over whatever torch.distributed backend is up (gloo on CPU in the tests,
   nccl on the GPU box); the library uses NCCL only to exchange the IPC handles and as a host-visible barrier.
"""
from __future__ import annotations

import numpy as np


def rows_per_rank(n: int, world: int, d: int) -> int:
    """Mirror of wb_comm_init: ceil(n / world) rounded up to whole block rows of the fused kernel and whole observation tiles, so
    every global row / tile of partial sums has exactly one writer."""
    import math
    v = (d + 3) // 4
    lanes = 1 if v <= 1 else 2 if v <= 2 else 4 if v <= 4 else 8
    pass_verts = 256 // lanes
    passes = -(-max(n, 1) // pass_verts)
    verts_per_block = pass_verts * max(1, -(-passes // (148 * 16)))
    align = math.lcm(verts_per_block, 1024)
    return -(-(-(-max(n, 1) // world)) // align) * align


def partition(n: int, world: int, d: int = 4) -> list[tuple[int, int]]:
    """[begin, end) of every rank (DeviceEmbedder.partition() returns the library's own answer)."""
    rows = rows_per_rank(n, world, d)
    return [(min(n, r * rows), min(n, r * rows + rows)) for r in range(world)]


def owner_of(v, n: int, world: int, d: int = 4):
    return np.asarray(v) // rows_per_rank(n, world, d)


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
    return dev.partition()
