"""``create_sequential_module`` (reference src/matten/model_factory/utils.py:13-91): threads the
output irreps of module i into the constructor of module i+1."""
from collections import OrderedDict
from typing import Dict, Optional, Tuple

from ..data.irreps import ModuleIrreps
from ..nn.sequential import Sequential
from ..o3 import Irreps


def create_sequential_module(modules: "OrderedDict[str, Tuple[ModuleIrreps, Dict]]",
                             irreps_in: Optional[Dict[str, Irreps]] = None,
                             use_kwargs_irreps_in: bool = False) -> Sequential:
    names, instances = [], []
    for name, (cls_type, kwargs) in modules.items():
        ir = irreps_in if not instances else instances[-1].irreps_out
        if "irreps_in" in kwargs:
            if not use_kwargs_irreps_in:
                raise ValueError(f"Trying to automatically determine irreps_in for module {name} But it is "
                                 f"provided as kwargs. Set `use_kwargs_irrpes_in=True` to force it.")
            if ir is not None:
                ir.update(kwargs["irreps_in"])
            else:
                ir = kwargs["irreps_in"]
        kwargs = dict(kwargs)
        kwargs["irreps_in"] = ir
        try:
            m = cls_type(**kwargs)
        except Exception as e:
            raise RuntimeError(f"Failed instantiate module `{cls_type.__name__}` with kwargs: `{kwargs}`") from e
        names.append(name)
        instances.append(m)
    return Sequential(OrderedDict(zip(names, instances)))
