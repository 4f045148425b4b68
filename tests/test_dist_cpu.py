"""-m "not gpu": the N>1 host logic on CPU with the gloo backend, world_size 2: contiguous frustum sharding
with no collective, the single flat gradient all-reduce of the data-parallel training steps, parameter broadcast and the
optional moving-statistics averaging."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from transferable3d_b200 import dist_util


def test_shard_range_partitions_contiguously():
    for total in (0, 1, 7, 32, 8192, 8191):
        for world in (1, 2, 3, 4, 8):
            edges = [dist_util.shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [e[1] - e[0] for e in edges]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        # per-replica gradients of a flat arena with per-variable views (like ParamArena)
        flat = torch.randn(1000, generator=g)
        views = [flat[:300], flat[300:301], flat[301:]]
        local = flat.clone()
        w = dist_util.allreduce_flat(flat)
        assert w == world
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        ref = torch.stack(gathered).sum(0)
        assert torch.allclose(flat, ref)
        assert torch.equal(views[1], flat[300:301])            # views alias the arena: nothing is packed / unpacked
        # parameters start identical
        p = torch.full((10,), float(rank))
        dist_util.broadcast_params(p, src=0)
        assert torch.equal(p, torch.zeros(10))
        # moving statistics averaging
        mv = [torch.full((4,), float(rank + 1)), torch.full((2, 3), float(10 * (rank + 1)))]
        dist_util.average_moving_stats(mv)
        assert torch.allclose(mv[0], torch.full((4,), 1.5)) and torch.allclose(mv[1], torch.full((2, 3), 15.0))
        # inference: shards cover the batch with no exchange; results of the ranks concatenate to the single-process result
        x = torch.arange(37, dtype=torch.float32)
        b, e = dist_util.shard_range(x.numel(), rank, world)
        y_local = x[b:e] * 2
        parts = [None] * world
        dist.all_gather_object(parts, y_local)
        assert torch.equal(torch.cat(parts), x * 2)
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_gloo_world2_flat_allreduce_and_sharding():
    world = 2
    ctx = mp.get_context('spawn')
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(out) == {0: 1, 1: 1}
