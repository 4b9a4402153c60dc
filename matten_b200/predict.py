"""
``predict()`` with the reference's signature (src/matten/predict.py:151-240) on the CUDA path.

Differences a caller can see:
  * structures may be pymatgen ``Structure`` objects (when pymatgen is installed) or plain dicts
    ``{"lattice": 3x3, "species": [Z or symbol, ...], "coords": n x 3, "coords_are_cartesian": bool}``;
  * elasticity tensors come back as ``pymatgen.analysis.elasticity.ElasticTensor`` only when pymatgen is
    importable, otherwise as ``numpy`` arrays of shape [3,3,3,3];
  * when ``torch.distributed`` is initialised with more than one rank, the structures are sharded over the ranks by
    edge count (no data-path collective; one all-gather of the [B, 21] results at the end) and every rank returns the
    full list.
Everything else (species check with the reference's message, ``failed_entries`` -> ``None``, flat per-atom list
for ``is_atomic_tensor``) follows the reference.
"""
from __future__ import annotations

import os
import pickle
import warnings
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from .model_factory.tfn_atomic_tensor import AtomicTensorModel
from .model_factory.tfn_scalar_tensor import ScalarTensorModel
from .nn.readout import CartesianTensorWrapper
from .parallel import gather_predictions, shard_by_edges

_SYMBOLS = ("H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr Rb Sr Y "
            "Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W "
            "Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn Fr Ra Ac Th Pa U Np Pu Am Cm Bk Cf Es Fm Md No Lr").split()
_Z_OF = {s: i + 1 for i, s in enumerate(_SYMBOLS)}


# ------------------------------------------------------------------------------------------ structures
def _structure_arrays(s) -> Dict[str, Any]:
    """(lattice [3,3], cartesian coords [n,3], atomic numbers [n]) of a pymatgen Structure or a plain dict."""
    if isinstance(s, dict):
        lat = np.asarray(s["lattice"]["matrix"] if isinstance(s["lattice"], dict) else s["lattice"], dtype=np.float64)
        if "atomic_numbers" in s:
            Z = [int(z) for z in s["atomic_numbers"]]
        elif "species" in s:
            Z = [int(z) if not isinstance(z, str) else _Z_OF[z] for z in s["species"]]
        else:  # pymatgen Structure.as_dict(): sites with species lists
            Z = [_Z_OF[site["species"][0]["element"]] for site in s["sites"]]
        if "cart_coords" in s:
            cart = np.asarray(s["cart_coords"], dtype=np.float64)
        elif "sites" in s and "coords" not in s:
            cart = np.asarray([site["xyz"] for site in s["sites"]], dtype=np.float64)
        else:
            c = np.asarray(s["coords"], dtype=np.float64)
            cart = c if s.get("coords_are_cartesian", False) else c @ lat
        return {"lattice": lat.reshape(3, 3), "cart": cart.reshape(-1, 3), "Z": Z}
    # duck-typed pymatgen Structure
    return {"lattice": np.asarray(s.lattice.matrix, dtype=np.float64),
            "cart": np.asarray(s.cart_coords, dtype=np.float64),
            "Z": [int(z) for z in s.atomic_numbers]}


def check_species(model, structures: Sequence[Dict[str, Any]]):
    """reference src/matten/predict.py:96-114"""
    supported = set(model.hparams["dataset_hparams"]["allowed_species"])
    for i, s in enumerate(structures):
        numbers = set(s["Z"])
        if not numbers.issubset(supported):
            bad = ", ".join(f"{_SYMBOLS[n - 1] if 0 < n <= len(_SYMBOLS) else '?'} ({n})"
                            for n in sorted(numbers - supported))
            raise RuntimeError(f"Cannot make predictions for structure {i}. It contains species {bad} not supported "
                               f"by the model. The model were trained with species {supported}.")


# ------------------------------------------------------------------------------------------ checkpoints
class _Stub:
    """Placeholder for classes of packages that are not installed (pytorch_lightning, torchmetrics, matten, ...)
    referenced inside a Lightning checkpoint; only tensors and plain containers are read from it."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"state": state})


class _TolerantUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        try:
            return super().find_class(module, name)
        except Exception:
            return type(name, (_Stub,), {"__module__": module})


class _TolerantPickle:
    Unpickler = _TolerantUnpickler
    __name__ = "tolerant_pickle"

    @staticmethod
    def load(f, **kw):
        return _TolerantUnpickler(f, **kw).load()


def load_checkpoint(path: Union[str, os.PathLike]) -> Dict[str, Any]:
    """Lightning ``.ckpt`` (or a plain ``{"state_dict", "hyper_parameters"}`` / state_dict file) -> dict with
    ``state_dict`` and ``hyper_parameters``; classes of packages that are not installed unpickle as stubs."""
    # Safe load first (tensors + plain containers only).  Lightning checkpoints pickle hyper-parameter objects of
    # packages that may be missing: those need the full unpickler -- which executes whatever the file contains, so
    # only load checkpoints you trust (same caveat as the reference's torch.load / load_from_checkpoint).
    try:
        ck = torch.load(path, map_location="cpu", weights_only=True)
    except Exception:
        try:
            ck = torch.load(path, map_location="cpu", weights_only=False)
        except Exception:
            ck = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_TolerantPickle)
    if "state_dict" not in ck:
        ck = {"state_dict": ck, "hyper_parameters": {}}
    return ck


def get_pretrained_model_dir(identifier: str) -> Path:
    p = Path(identifier)
    if p.exists() and p.is_dir():
        return p
    return Path(__file__).resolve().parent.parent / "pretrained" / identifier


def get_pretrained_config(identifier: str, config_filename: str = "config_final.yaml") -> Dict[str, Any]:
    import yaml

    with open(get_pretrained_model_dir(identifier) / config_filename) as f:
        return yaml.safe_load(f)


_MODEL_CACHE: Dict[Any, Any] = {}


def get_pretrained_model(identifier: str, checkpoint: str = "model_final.ckpt", model_class=ScalarTensorModel,
                         device=None):
    """``model_class.load_from_checkpoint`` of the reference: hyper-parameters and weights from the checkpoint (the
    backbone section of ``config_final.yaml`` is the fallback when the checkpoint carries no hyper-parameters).
    The loaded model (weights on the device, kernel plans built) is kept for the next call as long as the checkpoint
    file is unchanged: repeated ``predict()`` calls then cost the forward only."""
    directory = get_pretrained_model_dir(identifier)
    ck_path = directory / checkpoint
    try:
        st = os.stat(ck_path)
        key = (str(ck_path.resolve()), st.st_mtime_ns, st.st_size, model_class.__name__, str(device))
    except OSError:
        key = None
    if key is not None and key in _MODEL_CACHE:
        return _MODEL_CACHE[key]
    model = _load_pretrained_model(identifier, checkpoint, model_class, device)
    if key is not None:
        _MODEL_CACHE.clear()  # one model at a time: checkpoints are large
        _MODEL_CACHE[key] = model
    return model


def _load_pretrained_model(identifier: str, checkpoint: str, model_class, device):
    directory = get_pretrained_model_dir(identifier)
    ck = load_checkpoint(directory / checkpoint)
    hp = ck.get("hyper_parameters") or {}
    backbone = hp.get("backbone_hparams")
    dataset = hp.get("dataset_hparams")
    if backbone is None:
        backbone = get_pretrained_config(identifier)["model"]
    if dataset is None:
        raise RuntimeError("the checkpoint holds no dataset_hparams (allowed_species); cannot build the model")
    model = model_class(dict(backbone), dict(dataset))
    sd = {k: v for k, v in ck["state_dict"].items() if not k.startswith("metrics.")}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    missing = [k for k in missing]
    # e3nn modules store constant buffers (w3j tables, output masks) that this implementation regenerates, and EMPTY
    # placeholders wherever a weight is supplied from outside (`tp.tp.weight` of the uvu product with
    # internal_weights=False, `act.activation.mul.weight`, the `bias` of every bias-free o3.Linear)
    unexpected = [k for k in unexpected
                  if not any(t in k for t in ("_w3j", "output_mask", "num_batches_tracked")) and sd[k].numel() > 0]
    if missing or unexpected:
        raise RuntimeError(f"checkpoint does not match the model: missing {missing[:5]}, unexpected {unexpected[:5]}")
    if device is not None:
        model = model.to(device)
    return model.eval()


# ------------------------------------------------------------------------------------------ evaluation
def _edge_counts(batch) -> List[int]:
    """edges per crystal of a GPU-built batch (num_neigh summed per graph; tiny)."""
    from . import ops

    ptr32 = batch["ptr"].to(torch.int32)
    return [int(v) for v in ops.segment_reduce(batch["num_neigh"].reshape(-1, 1).contiguous(), ptr32, "sum").reshape(-1).cpu()]


def _build_batches(structs: Sequence[Dict[str, Any]], idx: Sequence[int], batch_size: int, r_cut: float, device, dtype):
    """Batched graphs of the structures ``idx`` built on the GPU (``data.neighbors.batch_from_structures``), in chunks
    of ``batch_size``.  A crystal without any edge inside ``r_cut`` cannot be converted (the reference raises "After
    eliminating self edges, no edges remain" and skips it, dataset/structure_scalar_tensor.py:357-362): it is reported
    in ``failed`` and left out.  Yields (batch, kept indices, edge counts)."""
    from .data.neighbors import batch_from_structures

    out, failed = [], []
    for i in range(0, len(idx), batch_size):
        chunk = list(idx[i:i + batch_size])
        batch = batch_from_structures([structs[j] for j in chunk], r_cut, device, dtype)
        counts = _edge_counts(batch)
        bad = [j for j, c in zip(chunk, counts) if c == 0]
        if bad:
            for j in bad:
                warnings.warn(f"Failed converting structure {j}: After eliminating self edges, no edges remain in "
                              f"this system. Skip it.")
            failed += bad
            chunk = [j for j in chunk if j not in set(bad)]
            if not chunk:
                continue
            batch = batch_from_structures([structs[j] for j in chunk], r_cut, device, dtype)
            counts = [c for c in counts if c > 0]
        out.append((batch, chunk, counts))
    return out, failed


def evaluate(model, batches, tensor_target_name: str, tensor_target_formula: str = "ijkl=jikl=klij") -> List[torch.Tensor]:
    """reference src/matten/predict.py:117-148: one forward per batch, irreps -> Cartesian."""
    converter = CartesianTensorWrapper(tensor_target_formula)
    out: List[torch.Tensor] = []
    model.eval()
    with torch.no_grad():
        for batch in batches:
            p = model(batch)[tensor_target_name]
            out.extend(converter.to_cartesian(p).cpu())
    return out


def predict(structure, model_identifier="20230627", checkpoint: str = "model_final.ckpt", batch_size: int = 200,
            logger_level: str = "ERROR", is_elasticity_tensor: bool = True, is_atomic_tensor: bool = False,
            device: Optional[Union[str, torch.device]] = None):
    """Predict the tensor property of a structure or a list of structures (see the module docstring)."""
    single = not isinstance(structure, (list, tuple))
    structs = [_structure_arrays(s) for s in ([structure] if single else structure)]
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("matten_b200.predict needs a CUDA device (there is no CPU fallback)")
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if is_atomic_tensor:
        model_class, is_elasticity_tensor = AtomicTensorModel, False
    else:
        model_class = ScalarTensorModel
    model = get_pretrained_model(model_identifier, checkpoint, model_class, device)
    check_species(model, structs)
    config = get_pretrained_config(model_identifier)
    name = config["data"]["tensor_target_name"]
    formula = config["data"]["tensor_target_formula"]
    r_cut = float(config["data"]["r_cut"])
    model.task_name = name
    dtype = next(model.parameters()).dtype
    rank_dims = (3,) * len(formula.split("=")[0])

    dist = torch.distributed
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    built, failed = _build_batches(structs, list(range(len(structs))), batch_size, r_cut, device, dtype)
    if world > 1 and not is_atomic_tensor:
        # shard the convertible crystals over the ranks by edge count; every rank ends up with the full result list
        kept = [j for _, chunk, _ in built for j in chunk]
        counts = [c for _, _, cs in built for c in cs]
        shards = shard_by_edges(counts, world)
        mine = [kept[i] for i in shards[dist.get_rank()]]
        del built
        mine_built, _ = _build_batches(structs, mine, batch_size, r_cut, device, dtype)
        local = evaluate(model, [b for b, _, _ in mine_built], name, formula)
        width = int(np.prod(rank_dims))
        flat = (torch.stack(local).reshape(len(local), -1) if local else torch.zeros((0, width), dtype=dtype)).to(device)
        allp = gather_predictions(flat, [len(s) for s in shards]).cpu()
        predictions = list(allp.reshape((allp.shape[0],) + rank_dims))
    else:
        predictions = evaluate(model, [b for b, _, _ in built], name, formula)

    if is_elasticity_tensor:
        try:
            from pymatgen.analysis.elasticity import ElasticTensor

            predictions = [ElasticTensor(t.numpy()) for t in predictions]
        except ImportError:
            predictions = [t.numpy() for t in predictions]
    else:
        predictions = [t.numpy() for t in predictions]

    if failed and not is_atomic_tensor:
        it = iter(predictions)
        fs = set(failed)
        predictions = [None if i in fs else next(it) for i in range(len(structs))]
        warnings.warn("Cannot make predictions for the following structures. Their returned elasticity tensor set "
                      f"to `None`: {sorted(fs)}.")
    if single and not is_atomic_tensor:
        return predictions[0]
    return predictions


def save_checkpoint(model, path: Union[str, os.PathLike]):
    """Write ``{"state_dict", "hyper_parameters"}`` in the layout of the reference's Lightning checkpoints (keys
    ``backbone.*`` / ``extra_layers_dict.*``), so that either implementation can load the other's weights."""
    torch.save({"state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()},
                "hyper_parameters": dict(model.hparams)}, path)


def save_pretrained(model, directory: Union[str, os.PathLike], r_cut: float, tensor_target_name: Optional[str] = None,
                    tensor_target_formula: Optional[str] = None, checkpoint: str = "model_final.ckpt"):
    """Writes a directory ``predict(model_identifier=directory)`` can use: the checkpoint plus ``config_final.yaml``
    with the ``data`` (r_cut, target name / formula) and ``model`` sections the reference's pretrained folders hold
    (pretrained/20230627/config_final.yaml)."""
    import yaml

    directory = Path(directory)
    directory.mkdir(parents=True, exist_ok=True)
    save_checkpoint(model, directory / checkpoint)
    hp = dict(model.hparams["backbone_hparams"])
    cfg = {"data": {"r_cut": float(r_cut), "tensor_target_name": tensor_target_name or model.task_name,
                    "tensor_target_formula": tensor_target_formula or hp.get("output_formula", "ijkl=jikl=klij")},
           "model": hp}
    with open(directory / "config_final.yaml", "w") as f:
        yaml.safe_dump(cfg, f)
    return directory
