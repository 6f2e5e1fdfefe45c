"""world_size-2 gloo tests of the frame-sharding host logic (runs on CPU)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepgraphpose_b200 import sharding
from oracle import dgp_ops


def test_shard_range_covers():
    for T in (0, 1, 7, 100, 10001):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(T, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == T
            for a, b in zip(r[:-1], r[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, T, nj, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    mu = torch.rand(T, nj, 2, generator=g) * 50
    a, b = sharding.shard_range(T, rank, world)
    local = mu[a:b]
    halo = sharding.exchange_halo(local[0])
    ext = torch.cat([local, halo[None]]) if halo is not None else local
    temporal_local = dgp_ops.temporal_distances(ext)      # rows a .. b-1 (last rank: a .. b-2)
    if halo is None:
        temporal_local = torch.cat([temporal_local, torch.zeros(1, nj)])
    full = sharding.gather_frames(temporal_local, T)[: T - 1]
    ref = dgp_ops.temporal_distances(mu)
    ok = torch.equal(full, ref)
    gathered = sharding.gather_frames(local, T)
    ok = ok and torch.equal(gathered, mu)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_halo_exchange_matches_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, T, nj = 2, 11, 3
    procs = [ctx.Process(target=_worker, args=(r, world, 29611, T, nj, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
