"""Point convolution (reference src/matten/nn/conv.py): self-connection + lin1 -> fused
[gather, radial MLP, uvu tensor product, segmented sum, 1/sqrt(#neigh)] -> lin2 -> add, and
the variant followed by Gate + BatchNorm.  Same constructor arguments, submodule names
(``lin1``, ``tp.weight_nn.layer*``, ``lin2``, ``sc``, ``act``, ``norm.n``) and parameter
layouts as the reference, so its checkpoints load."""
from typing import Dict

import torch

from ..data.irreps import DataKey, ModuleIrreps
from ..graph import get_graph
from ._nequip import with_batch
from ..o3 import Irreps
from .utils import ActivationLayer, NormalizationLayer, SpeciesLinear, UVUTensorProduct


class PointConv(ModuleIrreps, torch.nn.Module):
    def __init__(self, irreps_in: Dict[str, Irreps], conv_layer_irreps: Irreps, fc_num_hidden_layers: int = 1,
                 fc_hidden_size: int = 8, avg_num_neighbors: int = None):
        super().__init__()
        self.avg_num_neighbors = avg_num_neighbors
        self.init_irreps(irreps_in)
        x_ir = self.irreps_in[DataKey.NODE_FEATURES]
        attrs_ir = self.irreps_in[DataKey.NODE_ATTRS]
        sh_ir = self.irreps_in[DataKey.EDGE_ATTRS]
        conv_layer_irreps = Irreps(conv_layer_irreps)
        if len(attrs_ir) != 1 or attrs_ir[0].ir != (0, 1):
            raise NotImplementedError("node attributes must be a one-hot species encoding (Sx0e)")
        S = attrs_ir[0].mul
        self.num_species = S

        self.lin1 = SpeciesLinear(x_ir, S, x_ir)
        self.tp = UVUTensorProduct(x_ir, sh_ir, conv_layer_irreps,
                                   mlp_input_size=self.irreps_in[DataKey.EDGE_EMBEDDING].dim,
                                   mlp_hidden_size=fc_hidden_size, mlp_num_hidden_layers=fc_num_hidden_layers,
                                   mlp_activation="silu")
        self.lin2 = SpeciesLinear(self.tp.irreps_out, S, conv_layer_irreps)
        self.sc = SpeciesLinear(x_ir, S, conv_layer_irreps)
        self.irreps_out[DataKey.NODE_FEATURES] = conv_layer_irreps

    def forward(self, data: DataKey.Type) -> DataKey.Type:
        x = data[DataKey.NODE_FEATURES]
        g = get_graph(data)
        if self.num_species > 1:
            sperm, sptr = g.species_groups(data[DataKey.SPECIES_INDEX], self.num_species)
        else:
            sperm = sptr = None

        sc = self.sc(x, sperm, sptr)
        h = self.lin1(x, sperm, sptr)
        if self.avg_num_neighbors is not None:
            agg = self.tp.fused(h, data[DataKey.EDGE_ATTRS], data[DataKey.EDGE_EMBEDDING], g,
                                float(self.avg_num_neighbors))
        else:
            agg = self.tp.fused(h, data[DataKey.EDGE_ATTRS], data[DataKey.EDGE_EMBEDDING], g, None,
                                data[DataKey.NUM_NEIGH].reshape(-1).to(x.dtype))
        data[DataKey.NODE_FEATURES] = self.lin2(agg, sperm, sptr, residual=sc)  # sc + lin2(agg)
        return data


class PointConvWithActivation(ModuleIrreps, torch.nn.Module):
    def __init__(self, irreps_in: Dict[str, Irreps], conv_layer_irreps: Irreps, fc_num_hidden_layers: int = 1,
                 fc_hidden_size: int = 8, avg_num_neighbors: int = None, activation_type: str = "gate",
                 activation_scalars: Dict[str, str] = {"e": "silu", "o": "tanh"},
                 activation_gates: Dict[str, str] = {"e": "sigmoid", "o": "tanh"}, normalization: str = None):
        super().__init__()
        self.init_irreps(irreps_in)
        x_ir = self.irreps_in[DataKey.NODE_FEATURES]
        sh_ir = self.irreps_in[DataKey.EDGE_ATTRS]
        self.act = ActivationLayer(x_ir, sh_ir, Irreps(conv_layer_irreps), activation_type=activation_type,
                                   activation_scalars=activation_scalars, activation_gates=activation_gates)
        self.conv = PointConv(irreps_in=self.irreps_in, conv_layer_irreps=self.act.irreps_in,
                              fc_num_hidden_layers=fc_num_hidden_layers, fc_hidden_size=fc_hidden_size,
                              avg_num_neighbors=avg_num_neighbors)
        self.norm = NormalizationLayer(self.act.irreps_out, method=normalization)
        self.irreps_out[DataKey.NODE_FEATURES] = self.act.irreps_out

    def forward(self, data: DataKey.Type) -> DataKey.Type:
        data = self.conv(data)
        x = data[DataKey.NODE_FEATURES]
        if self.norm.method == "batch" and not self.norm.n.training:
            # eval: BatchNorm is a per-channel affine map -> folded into the gate kernel
            a, b = self.norm.n.eval_affine(x.dtype)
            x = self.act(x, a, b)
        else:
            x = self.act(x)
            if self.norm.method == "instance":
                with_batch(data)
                g = get_graph(data)
                x = self.norm(x, graph_ptr=g.graph_ptr(data[DataKey.BATCH], data.get("num_graphs")))
            else:
                x = self.norm(x, data.get(DataKey.BATCH))
        data[DataKey.NODE_FEATURES] = x
        return data
