"""Node-wise operations (reference src/matten/nn/nodewise.py)."""
from typing import Dict, Optional

import torch

from .. import functional as F
from ..data.irreps import DataKey, ModuleIrreps
from ..graph import get_graph
from ..o3 import Irreps
from ._nequip import with_batch
from .utils import IrrepsLinear


class NodewiseSelect(ModuleIrreps, torch.nn.Module):
    """Boolean-mask row selection (reference src/matten/nn/nodewise.py:18-86); pure indexing."""

    def __init__(self, irreps_in: Dict[str, Irreps], field: str = DataKey.NODE_FEATURES,
                 out_field: Optional[str] = None, mask_field: Optional[str] = None):
        super().__init__()
        self.field = field
        self.out_field = out_field if out_field is not None else field
        self.mask_field = mask_field
        self.init_irreps(irreps_in=irreps_in, irreps_out={self.out_field: irreps_in[self.field]},
                         required_keys_irreps_in=[self.field])

    def forward(self, data: DataKey.Type) -> DataKey.Type:
        data = data.copy()
        value = data[self.field]
        data[self.out_field] = value if self.mask_field is None else value[data[self.mask_field]]
        return data


class NodewiseLinear(ModuleIrreps, torch.nn.Module):
    """reference src/matten/nn/nodewise.py:89-117"""

    def __init__(self, irreps_in: Dict[str, Irreps], irreps_out: Irreps = None, field: str = DataKey.NODE_FEATURES,
                 out_field: Optional[str] = None):
        super().__init__()
        self.field = field
        self.out_field = out_field if out_field is not None else field
        if irreps_out is None:
            irreps_out = irreps_in[self.field]
        self.init_irreps(irreps_in=irreps_in, irreps_out={self.out_field: irreps_out},
                         required_keys_irreps_in=[self.field])
        self.linear = IrrepsLinear(self.irreps_in[field], self.irreps_out[self.out_field])

    def forward(self, data: DataKey.Type) -> DataKey.Type:
        data[self.out_field] = self.linear(data[self.field])
        return data


class NodewiseReduce(ModuleIrreps, torch.nn.Module):
    """Pooling over the nodes of each graph (reference src/matten/nn/nodewise.py:120-148): the
    batch vector is sorted, so this is a segmented reduction over graph pointers."""

    def __init__(self, irreps_in: Dict[str, Irreps], field: str, out_field: Optional[str] = None,
                 reduce: str = "sum"):
        super().__init__()
        assert reduce in ("sum", "mean", "min", "max")
        self.reduce, self.field = reduce, field
        self.out_field = f"{reduce}_{field}" if out_field is None else out_field
        self.init_irreps(irreps_in=irreps_in, irreps_out={self.out_field: irreps_in[self.field]},
                         required_keys_irreps_in=[self.field])

    def forward(self, data: DataKey.Type) -> DataKey.Type:
        with_batch(data)
        g = get_graph(data)
        ptr = g.graph_ptr(data[DataKey.BATCH], data.get("num_graphs"))
        data[self.out_field] = F.segment_reduce(data[self.field], ptr, self.reduce)
        return data
