"""N > 1 host logic on CPU: two gloo ranks agree on the partition and on the NCCL id shipped from rank 0."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from wembed_b200 import sharding


def test_partition_covers_every_vertex_once():
    for n, world in ((10, 2), (11, 4), (1_000_000, 8), (3, 8), (0, 2), (1, 1)):
        parts = sharding.partition(n, world)
        assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        rows = sharding.rows_per_rank(n, world, 4)
        assert rows % 1024 == 0 and rows * world >= n           # whole observation tiles and block rows per rank
        assert all(0 <= e - b <= rows for b, e in parts)
        assert all(b == n for b, e in parts if e == b and n)    # a rank without vertices sits at the end (wb_api.cu: it launches no tiles)
        if n:
            own = sharding.owner_of(np.arange(n), n, world)
            for r, (b, e) in enumerate(parts):
                assert (own[b:e] == r).all()


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = sharding.exchange_unique_id(lambda: bytes(range(128)), rank, world)
        lo, hi = sharding.partition(3001, world)[rank]
        # fixed-order sum of per-rank partials (every rank adds the same rows in the same order, as k_reduce_rows does on the device)
        import torch
        mine = torch.tensor([float(hi - lo), float(rank + 1) * 0.1], dtype=torch.float64)
        gathered = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, mine)
        total = sum(g for g in gathered)
        out.put((rank, uid, lo, hi, total.tolist()))
    finally:
        dist.destroy_process_group()


def test_two_ranks_share_id_and_partition():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, id0, lo0, hi0, t0), (r1, id1, lo1, hi1, t1) = res
    assert id0 == id1 == bytes(range(128))
    assert (lo0, hi0, lo1, hi1) == (0, 2048, 2048, 3001)
    assert t0 == t1 and t0[0] == 3001.0
