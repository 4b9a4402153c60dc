"""Synthetic crystals of BASELINE.json / SURVEY.md section 8d (config 2): 2x2x2 supercells of the
8-atom diamond-cubic cell (a = 5.43 A, 64 atoms, 10.86 A box) with N(0, 0.05 A) jitter; at
r_cut = 5 A every atom has exactly 28 neighbours (shells 2.35 / 3.84 / 4.50 A)."""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch

from .neighbors import collate, make_graph

DIAMOND_FRAC = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0],
                         [.25, .25, .25], [.25, .75, .75], [.75, .25, .75], [.75, .75, .25]])
SPECIES8 = [1, 6, 7, 8, 14, 22, 26, 29]


def diamond_supercell(rep: int = 2, a: float = 5.43):
    cells = np.stack(np.meshgrid(*[np.arange(rep)] * 3, indexing="ij"), -1).reshape(-1, 3)
    frac = (cells[:, None, :] + DIAMOND_FRAC[None]).reshape(-1, 3) / rep
    cell = np.eye(3) * a * rep
    return frac @ cell, cell


def synthetic_crystals(num: int, rep: int = 2, jitter: float = 0.05, species: Sequence[int] = SPECIES8,
                       r_cut: float = 5.0, seed: int = 0, dtype=torch.float32) -> List[Dict[str, torch.Tensor]]:
    gen = torch.Generator().manual_seed(seed)
    pos0, cell = diamond_supercell(rep)
    n = len(pos0)
    out = []
    sp = np.asarray(species)
    for _ in range(num):
        dp = torch.randn((n, 3), generator=gen, dtype=torch.float64).numpy() * jitter
        z = sp[torch.randint(0, len(sp), (n,), generator=gen).numpy()]
        out.append(make_graph(pos0 + dp, cell, z, r_cut, dtype=dtype))
    return out


def synthetic_batch(num: int, **kw) -> Dict[str, torch.Tensor]:
    return collate(synthetic_crystals(num, **kw))


def tile_batch(batch: Dict[str, torch.Tensor], times: int, jitter: float = 0.0, seed: int = 1):
    """Replicates a collated batch `times` times along the graph axis (fresh jitter on the copies keeps
    the topology): a cheap way to build the 512-crystal workload from a smaller neighbour search."""
    N = batch["pos"].shape[0]
    B = batch["num_graphs"]
    gen = torch.Generator().manual_seed(seed)
    out = {}
    pos = []
    for t in range(times):
        p = batch["pos"]
        if jitter > 0 and t > 0:
            p = p + torch.randn(p.shape, generator=gen, dtype=torch.float64).to(p.dtype) * jitter
        pos.append(p)
    out["pos"] = torch.cat(pos, 0)
    out["edge_index"] = torch.cat([batch["edge_index"] + t * N for t in range(times)], 1)
    out["batch"] = torch.cat([batch["batch"] + t * B for t in range(times)])
    for k in ("edge_cell_shift", "cell", "num_neigh", "atomic_numbers"):
        out[k] = torch.cat([batch[k]] * times, 0)
    out["num_graphs"] = B * times
    return out
