"""Per-atom tensor model (reference src/matten/model_factory/tfn_atomic_tensor.py): same backbone,
no pooling, the last NodewiseLinear maps straight onto ``CartesianTensor(formula)``."""
from typing import Any, Dict

from torch import Tensor

from ..nn.nodewise import NodewiseLinear
from ..nn.readout import ToCartesian
from ..o3 import CartesianTensor
from .tfn_scalar_tensor import OUT_FIELD_NAME, _common_layers, _TensorModelBase
from .utils import create_sequential_module


def create_model(hparams: Dict[str, Any], dataset_hparams: Dict[str, Any]):
    """reference src/matten/model_factory/tfn_atomic_tensor.py:103-199"""
    layers = _common_layers(hparams, dataset_hparams)
    formula = hparams["output_formula"].lower()
    layers["conv_to_output_hidden"] = (NodewiseLinear, {
        "irreps_out": CartesianTensor(formula=formula),
        "out_field": OUT_FIELD_NAME,
    })
    return create_sequential_module(modules=layers)


class AtomicTensorModel(_TensorModelBase):
    task_name = "nmr_tensor"

    def init_backbone(self, backbone_hparams, dataset_hparams):
        backbone = create_model(backbone_hparams, dataset_hparams)
        formula = backbone_hparams["output_formula"].lower()
        if backbone_hparams.get("output_format", "irreps") == "cartesian" and formula != "scalar":
            self.to_cartesian = ToCartesian(formula)
        else:
            self.to_cartesian = None
        return backbone, None

    def decode(self, model_input) -> Dict[str, Tensor]:
        out = self.backbone(model_input)[OUT_FIELD_NAME]
        if self.to_cartesian is not None:
            out = self.to_cartesian(out)
        return {self.task_name: out}
