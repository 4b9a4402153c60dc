"""
Multi-GPU decomposition: crystals are independent graphs, so inference shards the crystal
list across ranks with no data-path collective; training is data parallel (``matten_b200.train.Trainer``).
Host-side integer logic only -- testable on CPU with the gloo backend.
"""
from __future__ import annotations

from typing import List, Sequence

import torch


def shard_by_edges(edge_counts: Sequence[int], world_size: int) -> List[List[int]]:
    """Contiguous partition of crystals 0..n-1 into ``world_size`` shards with balanced EDGE counts (the conv cost
    is per edge, not per crystal).  Greedy prefix split at the ideal boundaries; every crystal lands in exactly one
    shard and the order is preserved, so concatenating the per-rank outputs restores the input order."""
    n = len(edge_counts)
    total = float(sum(edge_counts))
    shards: List[List[int]] = [[] for _ in range(world_size)]
    if n == 0:
        return shards
    acc = 0.0
    r = 0
    for i, c in enumerate(edge_counts):
        # move to the next shard when the midpoint of this crystal lies beyond the shard's ideal end
        while r < world_size - 1 and acc + 0.5 * c > total * (r + 1) / world_size:
            r += 1
        shards[r].append(i)
        acc += c
    return shards


def gather_predictions(local: torch.Tensor, counts: Sequence[int], group=None) -> torch.Tensor:
    """Concatenate the per-rank ``[B_k, D]`` outputs in rank order on every rank (control-plane collective at the END
    of inference; the data path itself has none)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    assert len(counts) == world
    D = local.shape[1]
    bufs = [torch.empty((c, D), dtype=local.dtype, device=local.device) for c in counts]
    if all(c == counts[0] for c in counts):
        dist.all_gather(bufs, local.contiguous(), group=group)
    else:
        mx = max(counts)
        pad = torch.zeros((mx, D), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
        tmp = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(tmp, pad, group=group)
        bufs = [t[:c] for t, c in zip(tmp, counts)]
    return torch.cat(bufs, 0)


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place mean of a flat gradient buffer over the data-parallel group (NCCL over NVLink on GPUs)."""
    import torch.distributed as dist

    dist.all_reduce(flat, group=group)
    flat.div_(dist.get_world_size(group))
    return flat
