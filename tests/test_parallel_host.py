"""Host-side logic of the multi-GPU decomposition on CPU: edge-balanced sharding, and the
rank exchange (result gather, gradient mean) over a world_size-2 gloo group -- the same code path NCCL takes on GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from matten_b200.parallel import allreduce_mean_, gather_predictions, shard_by_edges


def test_shard_by_edges_partitions_in_order_and_balances():
    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        for n in (0, 1, 7, 100, 513):
            counts = rng.integers(20, 4000, size=n).tolist()
            shards = shard_by_edges(counts, world)
            assert len(shards) == world
            flat = [i for s in shards for i in s]
            assert flat == list(range(n))  # every crystal exactly once, order preserved
            if n >= 8 * world:
                loads = [sum(counts[i] for i in s) for s in shards]
                assert max(loads) - min(loads) <= 2 * max(counts)  # within one crystal of the ideal split
    # uniform crystals split evenly
    assert [len(s) for s in shard_by_edges([1792] * 512, 8)] == [64] * 8


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # uneven shards of a [7, 21] result
        counts = [5, 2]
        full = torch.arange(7 * 21, dtype=torch.float32).reshape(7, 21)
        lo = sum(counts[:rank])
        got = gather_predictions(full[lo:lo + counts[rank]], counts)
        ok1 = torch.equal(got, full)
        # even shards take the plain all_gather path
        got2 = gather_predictions(full[3 * rank:3 * rank + 3], [3, 3])
        ok2 = torch.equal(got2, full[:6])
        g = torch.full((10,), float(rank + 1))
        allreduce_mean_(g)
        ok3 = torch.allclose(g, torch.full((10,), 1.5))
        q.put((rank, ok1, ok2, ok3))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gather_and_gradient_mean():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(all(r[1:]) for r in res), res
