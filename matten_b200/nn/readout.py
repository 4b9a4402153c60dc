"""Irreps -> Cartesian tensor (reference src/matten/nn/readout.py and src/matten/utils.py:110-133)."""
from typing import Dict, Optional

import torch

from .. import functional as F
from .. import o3, ops
from ..data.irreps import DataKey, ModuleIrreps
from ..plan import LinBlock


class CartesianTensorWrapper:
    """``CartesianTensor`` + cached change of basis (reference src/matten/utils.py:110-124).
    ``to_cartesian`` is ``v @ Q`` ([*,21] x [21,81] for the elasticity tensor) through the dense
    block of the linear kernel; ``from_cartesian`` is ``t @ Q^T``."""

    def __init__(self, formula: str):
        self.converter = o3.CartesianTensor(formula)
        self.formula = formula
        self.rank = len(self.converter.indices)
        self.dim = self.converter.dim
        self.ncart = 3 ** self.rank
        self._q = {}
        self._to = ops.LinPlanHandle([LinBlock(0, 0, self.dim, self.ncart, 1, 0, 1.0)], self.dim, self.ncart, 1,
                                     self.dim * self.ncart)
        self._from = ops.LinPlanHandle([LinBlock(0, 0, self.ncart, self.dim, 1, 0, 1.0)], self.ncart, self.dim, 1,
                                       self.dim * self.ncart)

    def _Q(self, t):
        key = (t.dtype, t.device)
        q = self._q.get(key)
        if q is None:
            Q = self.converter.change_of_basis(torch.float64)  # [dim, 3**rank]
            q = (Q.to(device=t.device, dtype=t.dtype).contiguous(),
                 Q.T.to(device=t.device, dtype=t.dtype).contiguous())
            self._q[key] = q
        return q

    def to_cartesian(self, data: torch.Tensor) -> torch.Tensor:
        Q, _ = self._Q(data)
        out = F.linear(self._to, data, Q)
        return out.reshape(tuple(data.shape[:-1]) + (3,) * self.rank)

    def from_cartesian(self, data: torch.Tensor) -> torch.Tensor:
        _, Qt = self._Q(data)
        return F.linear(self._from, data.flatten(-self.rank), Qt)


class ToCartesian(torch.nn.Module):
    """reference src/matten/utils.py:127-133"""

    def __init__(self, formula):
        super().__init__()
        self.ct = CartesianTensorWrapper(formula)

    def forward(self, data):
        return self.ct.to_cartesian(data)


class IrrepsToCartesianTensor(ModuleIrreps, torch.nn.Module):
    """reference src/matten/nn/readout.py:10-52"""

    def __init__(self, irreps_in: Dict[str, o3.Irreps], formula: str = "ij=ji", field: str = DataKey.NODE_FEATURES,
                 out_field: Optional[str] = None):
        super().__init__()
        self.formula, self.field = formula, field
        self.out_field = field if out_field is None else out_field
        self.init_irreps(irreps_in, required_keys_irreps_in=[field])
        self.ct = CartesianTensorWrapper(formula)
        assert self.irreps_in[self.field] == self.ct.converter, (
            f"input irreps of {self.field} is {self.irreps_in[self.field]}, not equal to the irreps of the "
            f"target irreps {self.ct.converter}")

    def forward(self, data: DataKey.Type) -> DataKey.Type:
        data[self.out_field] = self.ct.to_cartesian(data[self.field])
        return data
