"""Multi-GPU host logic on CPU (gloo, world_size 2): chunk tiles are dealt round-robin to ranks with no
data-path collective; the only communication is the gather of result descriptors / timings
that bench.py does with a barrier + all_reduce(MAX)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lizardfs_b200 import sharding


def test_tiles_cover_batch_exactly():
    for n, t in [(4096, 512), (4096, 500), (1, 8), (17, 4), (1024, 256)]:
        tl = sharding.tiles(n, t)
        assert sum(c for _, c in tl) == n and tl[0][0] == 0
        assert all(a[0] + a[1] == b[0] for a, b in zip(tl, tl[1:]))
        for w in (1, 2, 4, 8):
            seen = sorted(i for r in range(w) for i, _, _ in sharding.tiles_for_rank(n, t, r, w))
            assert seen == list(range(len(tl)))
            order = sharding.gather_order(n, t, w)
            assert len(order) == len(tl) and all(order[i][0] == i % w for i in range(len(tl)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_chunks, tile):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.tiles_for_rank(n_chunks, tile, rank, world)
    # each rank "encodes" its tiles: here the result descriptor is (tile index, first chunk, count)
    local = torch.tensor([[i, c0, n] for i, c0, n in mine] + [[-1, -1, -1]] * (64 - len(mine)), dtype=torch.int64)
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    # timing reduction used by bench.py: max over ranks
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == 10.0 + world - 1
    if rank == 0:
        rows = [tuple(r.tolist()) for g in gathered for r in g if r[0] >= 0]
        rows.sort()
        assert [r[0] for r in rows] == list(range(len(sharding.tiles(n_chunks, tile))))
        assert sum(r[2] for r in rows) == n_chunks
        order = sharding.gather_order(n_chunks, tile, world)
        for i, (owner, idx) in enumerate(order):
            assert tuple(gathered[owner][idx].tolist())[0] == i
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_two_ranks_gloo():
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 4096, 512), nprocs=2, join=True)
