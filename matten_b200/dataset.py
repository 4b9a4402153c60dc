"""
Reader of the reference's on-disk datasets (pandas-style column JSON, reference
src/matten/dataset/structure_scalar_tensor.py:230-367 + datasets/README.md): column ``structure`` holds pymatgen
``Structure.as_dict()`` records, the tensor target column (``elastic_tensor_full`` [3,3,3,3] per crystal, or
``nmr_tensor`` [n_selected,3,3] with ``atom_selector`` [n_atoms] per crystal) holds Cartesian tensors.  No pandas /
pymatgen needed: the records are parsed directly.  Targets are converted to irreps with the same
``CartesianTensor.from_cartesian`` projection as the reference (``tensor_target_format == "irreps"``), on the GPU.
"""
from __future__ import annotations

import json
from typing import Any, Dict, List, Optional

import numpy as np
import torch

from .nn.readout import CartesianTensorWrapper
from .predict import _structure_arrays


class TensorDataset:
    """structures (dicts with ``cart`` / ``lattice`` / ``Z``), targets (irreps or Cartesian) and, for per-atom targets,
    the selection mask of every crystal."""

    def __init__(self, filename: str, r_cut: float, tensor_target_name: str = "elastic_tensor_full",
                 tensor_target_format: str = "irreps", tensor_target_formula: str = "ijkl=jikl=klij",
                 atom_selector: Optional[str] = None, device="cuda", dtype=torch.float32):
        with open(filename) as f:
            cols = json.load(f)
        if "structure" not in cols:
            raise ValueError(f"Unsupported input data from file `{filename}`. Geometric information (e.g. pymatgen "
                             f"Structure) is needed, but the dataset does not have it.")
        keys = list(cols["structure"].keys())
        self.r_cut = float(r_cut)
        self.tensor_target_name = tensor_target_name
        self.structures: List[Dict[str, Any]] = [_structure_arrays(cols["structure"][k]) for k in keys]
        conv = CartesianTensorWrapper(tensor_target_formula)
        self.rank = conv.rank
        self.targets: List[torch.Tensor] = []
        self.selectors: Optional[List[torch.Tensor]] = [] if atom_selector else None
        dev = torch.device(device)
        for k in keys:
            t = torch.as_tensor(np.asarray(cols[tensor_target_name][k], dtype=np.float64)).to(device=dev, dtype=dtype)
            t = t.reshape((-1,) + (3,) * self.rank)
            if tensor_target_format == "irreps":
                t = conv.from_cartesian(t)  # symmetric projection, like the reference (NMR tensors are not symmetric)
            elif tensor_target_format != "cartesian":
                raise ValueError(f"Unsupported target tensor format `{tensor_target_format}`")
            self.targets.append(t)
            if atom_selector:
                self.selectors.append(torch.as_tensor(np.asarray(cols[atom_selector][k], dtype=bool)))
        self.failed_entries: List[Any] = []
        self._drop_edgeless(keys, dev)
        self.species = sorted({z for s in self.structures for z in s["Z"]})

    def _drop_edgeless(self, keys, dev, chunk: int = 256):
        """The reference skips crystals without any edge inside ``r_cut`` ("After eliminating self edges, no edges
        remain", dataset/structure_scalar_tensor.py:357-362) and records them; kept, they would divide by zero
        neighbours downstream."""
        from .data.neighbors import batch_from_structures
        from .predict import _edge_counts

        keep: List[int] = []
        for i in range(0, len(self.structures), chunk):
            b = batch_from_structures(self.structures[i:i + chunk], self.r_cut, dev, torch.float64)
            for j, c in enumerate(_edge_counts(b)):
                if c > 0:
                    keep.append(i + j)
                else:
                    self.failed_entries.append(keys[i + j])
        if self.failed_entries:
            import warnings

            warnings.warn(f"Skipped {len(self.failed_entries)} structures without any edge inside r_cut: "
                          f"{self.failed_entries[:10]}")
            self.structures = [self.structures[i] for i in keep]
            self.targets = [self.targets[i] for i in keep]
            if self.selectors is not None:
                self.selectors = [self.selectors[i] for i in keep]

    def __len__(self):
        return len(self.structures)

    def average_num_neighbors(self, device="cuda", batch_size: int = 256) -> float:
        """``avg_num_neighbors="auto"`` of the reference: mean of num_neigh over the (training) set
        (dataset/structure_scalar_tensor.py:659-663)."""
        from .data.neighbors import batch_from_structures

        tot, n = 0.0, 0
        for i in range(0, len(self), batch_size):
            b = batch_from_structures(self.structures[i:i + batch_size], self.r_cut, device, torch.float64)
            tot += float(b["num_neigh"].sum())
            n += int(b["num_neigh"].numel())
        return tot / max(n, 1)

    def batches(self, batch_size: int, device="cuda", dtype=torch.float32, shuffle: bool = False, seed: int = 0,
                rank: int = 0, world: int = 1, even: bool = True):
        """Yields (graph batch, target [*, dim], atom selector | None).  Under data parallelism every rank takes
        every ``world``-th batch.  ``even=True`` (training) drops the last chunks so that every rank runs the same
        number of steps (the gradient all-reduce needs matching steps); evaluation passes ``even=False`` and sees
        every crystal (its single all-reduce at the end tolerates uneven counts)."""
        from .data.neighbors import batch_from_structures

        order = np.arange(len(self))
        if shuffle:
            np.random.default_rng(seed).shuffle(order)
        chunks = [order[i:i + batch_size] for i in range(0, len(order), batch_size)]
        if world > 1 and even:
            chunks = chunks[: len(chunks) // world * world]
        for c in chunks[rank::world]:
            batch = batch_from_structures([self.structures[i] for i in c], self.r_cut, device, dtype)
            target = torch.cat([self.targets[i] for i in c], 0).to(dtype)
            sel = torch.cat([self.selectors[i] for i in c]).to(device) if self.selectors is not None else None
            yield batch, target, sel
