"""In/out irreps protocol of a module: the reference's ``ModuleIrreps``
(src/matten/data/irreps.py:17-209) on top of :mod:`matten_b200.o3`, so that
``create_sequential_module`` can thread ``irreps_out -> irreps_in`` unchanged."""
from typing import Dict, Optional, Sequence

from ..o3 import Irreps
from . import _key

DataKey = _key


def _fix_irreps_dict(d: Dict[str, Irreps]) -> Dict[str, Optional[Irreps]]:
    return {k: (None if v is None else Irreps(v)) for k, v in d.items()}


def _check_irreps_compatible(ir1: Dict[str, Irreps], ir2: Dict[str, Irreps]) -> bool:
    return all(ir1[k] == ir2[k] for k in ir1 if k in ir2)


def _check_irreps_type(irreps1, irreps2) -> bool:
    return set(Irreps(irreps1).ls) == set(Irreps(irreps2).ls)


class ModuleIrreps:
    REQUIRED_KEYS_IRREPS_IN = None
    REQUIRED_TYPE_IRREPS_IN = None
    OPTIONAL_MUL_TYPE_IRREPS_IN = None

    def init_irreps(self, irreps_in=None, irreps_out=None, *, required_keys_irreps_in: Sequence[str] = None,
                    required_type_irreps_in=None, optional_mul_type_irreps_in=None):
        irreps_in = self.fix_irreps_in(_fix_irreps_dict({} if irreps_in is None else irreps_in))
        if irreps_out is None:
            irreps_out = {}
        elif isinstance(irreps_out, str):
            assert irreps_out in irreps_in, f"`irreps_in` does not contain key for `irreps_out = {irreps_out}`"
            irreps_out = {irreps_out: irreps_in[irreps_out]}
        irreps_out = _fix_irreps_dict(irreps_out)

        required_keys = list(self.REQUIRED_KEYS_IRREPS_IN or [])
        if required_keys_irreps_in is not None:
            required_keys += list(required_keys_irreps_in)
        required_type = dict(self.REQUIRED_TYPE_IRREPS_IN or {})
        if required_type_irreps_in is not None:
            required_type.update(required_type_irreps_in)
        required_type = _fix_irreps_dict(required_type)
        optional = dict(self.OPTIONAL_MUL_TYPE_IRREPS_IN or {})
        if optional_mul_type_irreps_in is not None:
            optional.update(optional_mul_type_irreps_in)
        optional = _fix_irreps_dict(optional)

        for k in required_keys + list(required_type.keys()):
            if k not in irreps_in:
                raise ValueError(f"This module {type(self)} requires `{k}` in `irreps_in`.")
        for k, v in required_type.items():
            if not _check_irreps_type(irreps_in[k], v):
                raise ValueError(f"This module {type(self)} expects irreps_in['{k}'] be of type {v}, instead got "
                                 f"{irreps_in[k]}. Note, type means degree and parity, not multiplicity.")
        for k, v in optional.items():
            if k in irreps_in and irreps_in[k] != v:
                raise ValueError(f"This module {type(self)} expects irreps_in['{k}'] to be {v}, instead got "
                                 f"{irreps_in[k]}.")
        self._irreps_in = irreps_in
        self._irreps_out = irreps_in.copy()
        self._irreps_out.update(irreps_out)

    @property
    def irreps_in(self):
        return self._irreps_in

    @property
    def irreps_out(self):
        return self._irreps_out

    def fix_irreps_in(self, irreps_in):
        irreps_in = irreps_in.copy()
        pos = DataKey.POSITIONS
        if pos in irreps_in and irreps_in[pos] != Irreps("1x1o"):
            raise ValueError(f"Positions must have irreps 1o, got `{irreps_in[pos]}`")
        irreps_in[pos] = Irreps("1o")
        ei = DataKey.EDGE_INDEX
        if ei in irreps_in and irreps_in[ei] is not None:
            raise ValueError(f"Edge indexes must have irreps `None`, got `{irreps_in[ei]}`")
        irreps_in[ei] = None
        return irreps_in
