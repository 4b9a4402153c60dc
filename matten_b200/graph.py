"""
Per-batch index bookkeeping shared by all layers of a forward pass.

The reference scatter-adds messages with torch_scatter atomics (src/matten/nn/conv.py:114)
and multiplies by dense one-hot species attributes (conv.py:109-123).  Here the batch is
indexed ONCE: edges are stably sorted by receiver into a CSR (so the scatter is a
deterministic segmented sum), nodes are grouped by species (so the species-indexed
linears see one dense weight slice per tile) and the batch vector becomes graph pointers.
All of it is integer work on the GPU (matten_b200/csrc/graph_ops.cu), bit-exact against
numpy's stable argsort (tests/test_gpu_ops.py).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import ops
from .data import _key as K


class GraphCache:
    def __init__(self, data: Dict[str, torch.Tensor]):
        ei = data[K.EDGE_INDEX]
        self.device = ei.device
        self.N = int(data[K.POSITIONS].shape[0]) if K.POSITIONS in data else int(data[K.NODE_FEATURES].shape[0])
        self.E = int(ei.shape[1])
        self.flag = ops.new_flag(self.device)
        ei = ei if ei.is_contiguous() else ei.contiguous()
        self.edge_index = ei
        # receiver-sorted CSR; messages flow edge_index[0] -> edge_index[1] (conv.py:107-114)
        self.rowptr, self.perm = ops.csr_by_key(ei[1], self.N, True, self.flag)
        self.src_sorted = ops.gather_i64_to_i32(ei[0], self.perm)
        self._species = None
        self._graph_ptr = None
        self._sender = None
        self._conv_layout = None

    def species_groups(self, species_index: torch.Tensor, num_species: int):
        if self._species is None or self._species[0] != num_species:
            ptr, perm = ops.csr_by_key(species_index, num_species, True, self.flag)
            self._species = (num_species, perm, ptr)
        return self._species[1], self._species[2]

    def graph_ptr(self, batch: torch.Tensor, num_graphs: Optional[int] = None):
        if self._graph_ptr is None:
            if num_graphs is None:
                # batch is sorted, so the last entry is the largest graph id (one scalar read)
                num_graphs = int(batch[-1].item()) + 1 if batch.numel() else 0
            ops.check_sorted(batch, self.flag)
            ptr, _ = ops.csr_by_key(batch, num_graphs, False, self.flag)
            self._graph_ptr = ptr
        return self._graph_ptr

    def conv_layout(self, sh: torch.Tensor, y_lmax: Optional[int]):
        """Layer-invariant inputs of the fp32 tensor-core convolution, built by the first PointConv layer of the
        forward and reused by the others (all layers see the same graph and the same edge_attrs tensor)."""
        if sh.dtype != torch.float32 or y_lmax is None or self.E == 0:
            return None
        key = (sh.data_ptr(), sh._version, int(y_lmax))
        if self._conv_layout is None or self._conv_layout[0] != key:
            self._conv_layout = (key, ops.conv_layout(sh, y_lmax, self.rowptr, self.perm, self.src_sorted, self.N), sh)
        return self._conv_layout[1]

    def sender_csr(self):
        """CSR over senders of the receiver-sorted edge list (for the backward pass)."""
        if self._sender is None:
            keys = self.src_sorted.to(torch.int64)
            ptr, perm = ops.csr_by_key(keys, self.N, True, self.flag)
            self._sender = (ptr, perm)
        return self._sender

    def raise_if_invalid(self):
        ops.raise_on_flag(self.flag)


def get_graph(data: Dict[str, torch.Tensor]) -> GraphCache:
    g = data.get(K.GRAPH_CACHE)
    if g is None:
        g = GraphCache(data)
        data[K.GRAPH_CACHE] = g
    return g
