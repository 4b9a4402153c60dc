"""Graph-level tensor model (reference src/matten/model_factory/tfn_scalar_tensor.py): backbone
``Sequential`` + ``out_layer`` Linear after pooling + optional conversion to Cartesian.  The
Lightning shell of the reference (loss logging, callbacks) is out of scope; the arithmetic of
``decode`` / ``forward`` is identical and the ``state_dict`` keys (``backbone.*``,
``extra_layers_dict.out_layer.weight``) are the reference's."""
from collections import OrderedDict
from typing import Any, Dict, Optional

import torch
from torch import Tensor

from ..data import _key as K
from ..nn._nequip import SphericalHarmonicEdgeAttrs
from ..nn.conv import PointConv, PointConvWithActivation
from ..nn.embedding import EdgeLengthEmbedding, SpeciesEmbedding
from ..nn.nodewise import NodewiseLinear, NodewiseReduce
from ..nn.readout import ToCartesian
from ..nn.utils import IrrepsLinear
from ..o3 import CartesianTensor, Irreps
from .utils import create_sequential_module

OUT_FIELD_NAME = "my_model_output"


def _common_layers(hparams: Dict[str, Any], dataset_hparams: Dict[str, Any]) -> "OrderedDict":
    layers = OrderedDict()
    layers["one_hot"] = (SpeciesEmbedding, {
        "embedding_dim": hparams["species_embedding_dim"],
        "allowed_species": dataset_hparams["allowed_species"],
        "use_atom_feats": hparams.get("use_atom_feats", False),
        "atom_feats_dim": dataset_hparams.get("atom_feats_size", None),
    })
    layers["spharm_edges"] = (SphericalHarmonicEdgeAttrs, {"irreps_edge_sh": hparams["irreps_edge_sh"]})
    layers["radial_basis"] = (EdgeLengthEmbedding, {
        "num_basis": hparams["num_radial_basis"],
        "start": hparams["radial_basis_start"],
        "end": hparams["radial_basis_end"],
        "basis": hparams["radial_basis_type"],
    })
    num_neigh = hparams["average_num_neighbors"]
    if isinstance(num_neigh, str) and num_neigh.lower() == "auto":
        num_neigh = dataset_hparams["average_num_neighbors"]
    for i in range(hparams["num_layers"]):
        layers[f"layer{i}_convnet"] = (PointConvWithActivation, {
            "conv_layer_irreps": hparams["conv_layer_irreps"],
            "activation_type": hparams["nonlinearity_type"],
            "fc_num_hidden_layers": hparams["invariant_layers"],
            "fc_hidden_size": hparams["invariant_neurons"],
            "avg_num_neighbors": num_neigh,
            "normalization": hparams["normalization"],
        })
    layers["conv_layer_last"] = (PointConv, {
        "conv_layer_irreps": hparams["conv_layer_irreps"],
        "fc_num_hidden_layers": hparams["invariant_layers"],
        "fc_hidden_size": hparams["invariant_neurons"],
        "avg_num_neighbors": num_neigh,
    })
    return layers


def create_model(hparams: Dict[str, Any], dataset_hparams: Dict[str, Any]):
    """reference src/matten/model_factory/tfn_scalar_tensor.py:103-195"""
    layers = _common_layers(hparams, dataset_hparams)
    layers["conv_to_output_hidden"] = (NodewiseLinear, {
        "irreps_out": hparams["conv_to_output_hidden_irreps_out"],
        "out_field": OUT_FIELD_NAME,
    })
    layers["output_pooling"] = (NodewiseReduce, {
        "field": OUT_FIELD_NAME,
        "out_field": OUT_FIELD_NAME,
        "reduce": hparams["reduce"],
    })
    return create_sequential_module(modules=layers)


class _TensorModelBase(torch.nn.Module):
    task_name = "elastic_tensor_full"

    def __init__(self, backbone_hparams: Dict[str, Any], dataset_hparams: Optional[Dict[str, Any]] = None,
                 task_name: Optional[str] = None):
        super().__init__()
        self.hparams = {"backbone_hparams": dict(backbone_hparams), "dataset_hparams": dict(dataset_hparams or {})}
        if task_name is not None:
            self.task_name = task_name
        self.backbone, extra = self.init_backbone(backbone_hparams, dataset_hparams or {})
        self.extra_layers_dict = torch.nn.ModuleDict(extra) if extra is not None else None

    def preprocess(self, data: Dict[str, Tensor]) -> Dict[str, Tensor]:
        """Shallow copy of the graph dict (modules add keys to it), like the reference's
        ``tensor_property_to_dict`` output being consumed once per forward."""
        return dict(data)

    def forward(self, data: Dict[str, Tensor], check: bool = True, mode: Optional[str] = None,
                task_name: Optional[str] = None, return_labels: bool = False):
        """reference BaseModel.forward (src/matten/model/model.py:143-184): decode (+ identity
        target transform; the shipped configs use no normaliser).  ``mode="backbone"`` returns the backbone's graph
        dict, as in the reference.  DEVIATION: the reference takes a PyG ``DataPoint`` batch and returns the tuple
        ``(preds, labels)``; here the input is the batched graph dict and the default return is the ``preds`` dict
        alone -- pass ``return_labels=True`` for the reference's tuple (labels = ``data["y"]`` when present)."""
        if task_name is not None:
            self.task_name = task_name
        d = self.preprocess(data)
        if mode is not None and str(mode).lower() not in ("none", "backbone"):
            raise ValueError(f"Expect mode to be one of (None, 'backbone'); got {mode}")
        preds = self.backbone(d) if mode == "backbone" else self.decode(d)
        self._last_graph = d.get(K.GRAPH_CACHE)  # index bookkeeping + device error word of this forward
        if check and K.GRAPH_CACHE in d:
            d[K.GRAPH_CACHE].raise_if_invalid()
        if return_labels:
            return preds, data.get("y", {})
        return preds


class ScalarTensorModel(_TensorModelBase):
    def init_backbone(self, backbone_hparams, dataset_hparams):
        backbone = create_model(backbone_hparams, dataset_hparams)
        formula = backbone_hparams["output_formula"].lower()
        irreps_out = Irreps("0e") if formula == "scalar" else CartesianTensor(formula=formula)
        irreps_in = backbone_hparams["conv_to_output_hidden_irreps_out"]
        extra = {"out_layer": IrrepsLinear(irreps_in, irreps_out)}
        if backbone_hparams.get("output_format", "irreps") == "cartesian" and formula != "scalar":
            self.to_cartesian = ToCartesian(formula)
        else:
            self.to_cartesian = None
        return backbone, extra

    def decode(self, model_input) -> Dict[str, Tensor]:
        out = self.backbone(model_input)[OUT_FIELD_NAME]
        out = self.extra_layers_dict["out_layer"](out)
        if self.to_cartesian is not None:
            out = self.to_cartesian(out)
        return {self.task_name: out}
