"""N > 1 host logic on CPU: world_size-2 and -3 gloo groups shard a batch and gather the verdicts."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from numbotics_b200.dist import shard_bounds, gather_shards, max_over_ranks, sum_over_ranks


def test_shard_bounds_partition():
    for n in (0, 1, 7, 1000, 1 << 20, 1000003):
        for world in (1, 2, 3, 4, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _verdict_of(idx):
    # stands in for the per-rank kernel output: a deterministic function of the GLOBAL index
    return (idx * 2654435761 % 7) < 3


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(n, rank, world)
        idx = torch.arange(lo, hi)
        local = _verdict_of(idx)                                  # bool verdicts
        full = gather_shards(local, n)
        ok = torch.equal(full, _verdict_of(torch.arange(n)))
        poses = torch.stack([idx.double(), idx.double() * 2], dim=1)   # a (n,2) float result
        on0 = gather_shards(poses, n, dst=0)
        if rank == 0:
            ok = ok and torch.equal(on0[:, 0], torch.arange(n).double()) and on0.shape == (n, 2)
        else:
            ok = ok and on0 is None
        ok = ok and max_over_ranks(float(rank + 1)) == float(world)
        ok = ok and sum_over_ranks(float(hi - lo)) == float(n)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 1001), (3, 1000), (2, 4)])
def test_gather_shards_gloo(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    results = dict(q.get(timeout=10) for _ in range(world))
    assert results == {r: True for r in range(world)}


def test_single_process_gather_is_identity():
    t = torch.arange(5)
    assert gather_shards(t, 5) is t
    with pytest.raises(ValueError):
        gather_shards(t, 6)
