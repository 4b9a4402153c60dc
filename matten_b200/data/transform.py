"""Target normalisers (reference src/matten/data/transform.py:23-216 ``Normalize`` / ``MeanNormNormalize``,
:219-302 ``ScalarNormalize``, :417-619 the target transforms that own them).

The reference applies these per crystal on the CPU as a dataset ``pre_transform`` and inverts them on the model
output before the metrics (src/matten/model_factory/task.py:94-107).  Here both directions run on the batched
``[B, D]`` target / prediction matrix on the device (``mt_normalize``), and the dataset statistics are column
reductions over the ``[num_crystals, D]`` target matrix (``mt_col_reduce``), so a training step never leaves the GPU.
Same class names, constructor arguments, ``mean`` / ``norm`` buffers (state-dict compatible with the reference's
``dataset_statistics.pt`` entries) and error behaviour."""
from pathlib import Path
from typing import Dict, List, Optional, Union

import torch
import torch.nn as nn

from .. import ops
from ..o3 import Irreps


def _channel_matrix(irreps: Irreps, dtype, device):
    """[D, num_channels] 0/1 matrix: column j belongs to channel c; plus per-channel (2l+1, is 0e) lists."""
    D = irreps.dim
    nch = sum(mul for mul, _ in irreps)
    m = torch.zeros((D, nch), dtype=dtype, device=device)
    dims, scalar = [], []
    col = ch = 0
    for mul, ir in irreps:
        for _ in range(mul):
            m[col:col + ir.dim, ch] = 1
            dims.append(ir.dim)
            scalar.append(ir.is_scalar())
            col += ir.dim
            ch += 1
    return m, torch.tensor(dims, dtype=dtype, device=device), torch.tensor(scalar, dtype=torch.bool, device=device)


class Normalize(nn.Module):
    """Base class for tensor standardization (reference transform.py:23-56)."""

    def __init__(self, irreps: Union[str, Irreps]):
        super().__init__()
        self.irreps = Irreps(irreps)

    def forward(self, data):
        raise NotImplementedError

    def inverse(self, data):
        raise NotImplementedError


class _MeanNorm(nn.Module):
    """The shared ``mean`` / ``norm`` buffers and the two elementwise maps."""

    def _init_buffers(self, dim: int, mean, norm, scale: float):
        self.scale = scale
        # as in the reference: buffers are always registered, a flag says whether they hold statistics
        self.mean_norm_initialized = not (mean is None or norm is None)
        self.register_buffer("mean", torch.zeros(dim) if mean is None else mean)
        self.register_buffer("norm", torch.zeros(dim) if norm is None else norm)

    def _stats(self, data):
        if not self.mean_norm_initialized:
            raise RuntimeError("mean and norm not initialized.")
        if self.mean.device != data.device or self.mean.dtype != data.dtype:
            return self.mean.to(data.device, data.dtype), self.norm.to(data.device, data.dtype)
        return self.mean, self.norm

    def forward(self, data: torch.Tensor) -> torch.Tensor:
        mean, norm = self._stats(data)
        return ops.normalize(data, mean, norm, self.scale, inverse=False)

    def inverse(self, data: torch.Tensor) -> torch.Tensor:
        mean, norm = self._stats(data)
        return ops.normalize(data, mean, norm, self.scale, inverse=True)

    def load_state_dict(self, state_dict, strict: bool = True):
        # the buffers take the dtype of the statistics (the reference copies into its default-dtype buffers; its
        # data are default dtype too), so fp64 targets keep fp64 statistics
        dev = self.mean.device
        self.mean = state_dict["mean"].detach().clone().to(dev)
        self.norm = state_dict["norm"].detach().clone().to(dev)
        out = super().load_state_dict(state_dict, strict)
        self.mean_norm_initialized = True
        return out


class MeanNormNormalize(_MeanNorm, Normalize):
    """Normalise like e3nn BatchNorm (reference transform.py:59-216): 0e channels are centred, every channel is
    divided by the root of its mean squared (component-averaged or summed) norm over the dataset."""

    def __init__(self, irreps: Union[str, Irreps], mean=None, norm=None, normalization: str = "component",
                 reduce: str = "mean", eps: float = 1e-5, scale: float = 1.0):
        Normalize.__init__(self, irreps)
        self.normalization, self.reduce, self.eps = normalization, reduce, eps
        self._init_buffers(self.irreps.dim, mean, norm, scale)

    def compute_statistics(self, data: torch.Tensor):
        """``data`` [num_samples, D] on the device.  Returns (mean [D], norm [D]) and stores them."""
        if self.normalization not in ("norm", "component"):
            raise ValueError(f"Invalid normalization option {self.normalization}")
        if self.reduce != "mean":
            # the reference's `max` branch adds eps to the (values, indices) pair torch.max returns and fails
            # (transform.py:186-193); only `mean` is usable there
            raise ValueError("Invalid reduce option {}".format(self.reduce))
        dim = data.shape[-1]
        assert dim == self.irreps.dim, (f"`ix` should have reached data.size(-1)={dim}, but it ended at "
                                        f"{self.irreps.dim}")
        n = data.shape[0]
        chan, cdim, is_scalar = _channel_matrix(self.irreps, data.dtype, data.device)
        col_scalar = (chan @ is_scalar.to(data.dtype)) > 0
        colsum = ops.col_reduce(data)
        mean = torch.where(col_scalar, colsum / n, torch.zeros_like(colsum)).contiguous()
        ssq = ops.col_reduce(data, mean, data, mean)  # sum_n (x - mean)^2 per column
        per_chan = ssq @ chan  # summed over the 2l+1 components
        if self.normalization == "component":
            per_chan = per_chan / cdim
        norm_c = (per_chan / n + self.eps).pow(0.5)
        norm = chan @ norm_c  # expand back to columns
        self.load_state_dict({"mean": mean, "norm": norm})
        return mean, norm


class ScalarNormalize(_MeanNorm):
    """Per-feature standardisation of [num_samples, num_features] scalars (reference transform.py:219-302)."""

    def __init__(self, num_features: int, mean=None, norm=None, scale: float = 1.0):
        nn.Module.__init__(self)
        self._init_buffers(num_features, mean, norm, scale)

    def compute_statistics(self, data: torch.Tensor):
        """sklearn ``StandardScaler().fit``: column mean and population standard deviation, a zero deviation
        replaced by 1 (reference transform.py:281-302)."""
        assert data.ndim == 2, "Can only deal with tensor [N_samples, N_features]"
        n = data.shape[0]
        mean = (ops.col_reduce(data) / n).contiguous()
        var = ops.col_reduce(data, mean, data, mean) / n
        std = var.sqrt()
        std = torch.where(std < 10 * torch.finfo(data.dtype).eps, torch.ones_like(std), std)
        self.load_state_dict({"mean": mean, "norm": std})
        return mean, std


class _TargetTransform(nn.Module):
    def __init__(self, dataset_statistics_path: Union[str, Path, None]):
        super().__init__()
        self.dataset_statistics_path = dataset_statistics_path
        self.dataset_statistics_loaded = False

    def _load(self):
        if self.dataset_statistics_path is None:
            raise ValueError("Cannot load dataset statistics from file `None`")
        try:
            return torch.load(self.dataset_statistics_path, weights_only=True)
        except Exception:  # statistics files written by the reference hold plain containers and tensors only
            return torch.load(self.dataset_statistics_path, weights_only=False)


class TensorTargetTransform(_TargetTransform):
    """Forward / inverse normalisation of the tensor target (reference transform.py:527-619).  ``forward`` and
    ``inverse`` take the batched ``[B, D]`` irreps-layout targets; the statistics are loaded lazily from
    ``dataset_statistics_path`` on first use, as in the reference."""

    def __init__(self, target_name: str = "elastic_tensor_full", dataset_statistics_path: Union[str, Path] = None,
                 scale: float = 1.0, irreps: str = "2x0e+2x2e+4e"):
        super().__init__(dataset_statistics_path)
        self.target_name = target_name
        self.normalizer = MeanNormNormalize(irreps=irreps, scale=scale)

    def _fill_state_dict(self, device):
        if not self.dataset_statistics_loaded and not self.normalizer.mean_norm_initialized:
            self.normalizer.load_state_dict(self._load()[self.target_name])
        self.to(device)
        self.dataset_statistics_loaded = True

    def forward(self, target: torch.Tensor) -> torch.Tensor:
        self._fill_state_dict(target.device)
        return self.normalizer(target)

    def inverse(self, data: torch.Tensor) -> torch.Tensor:
        self._fill_state_dict(data.device)
        return self.normalizer.inverse(data)

    def compute_statistics(self, targets: torch.Tensor, atomic_numbers=None, num_neigh=None) -> Dict:
        """``targets`` [num_crystals, D]; returns the dictionary the reference saves as ``dataset_statistics.pt``."""
        self.normalizer.compute_statistics(targets)
        self.dataset_statistics_loaded = True
        stats = {self.target_name: {k: v.cpu() for k, v in self.normalizer.state_dict().items()}}
        if atomic_numbers is not None:
            stats["allowed_species"] = tuple(sorted(set(int(z) for z in atomic_numbers)))
        if num_neigh is not None:
            stats["average_num_neigh"] = torch.as_tensor(num_neigh, dtype=torch.get_default_dtype()).mean()
        return stats


class ScalarTargetTransform(_TargetTransform):
    """Forward / inverse normalisation of named scalar targets (reference transform.py:417-524)."""

    def __init__(self, target_names: List[str], dataset_statistics_path: Union[str, Path] = None):
        super().__init__(dataset_statistics_path)
        self.target_names = list(target_names)
        self.normalizers = nn.ModuleDict({name: ScalarNormalize(num_features=1) for name in self.target_names})

    def _fill_state_dict(self, device):
        if not self.dataset_statistics_loaded:
            pending = [n for n in self.target_names if not self.normalizers[n].mean_norm_initialized]
            if pending:
                stats = self._load()
                for name in pending:
                    self.normalizers[name].load_state_dict(stats[name])
        self.to(device)
        self.dataset_statistics_loaded = True

    def forward(self, targets: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        self._fill_state_dict(next(iter(targets.values())).device)
        for name in self.target_names:
            targets[name] = self.normalizers[name](targets[name])
        return targets

    def inverse(self, data: torch.Tensor, target_name: str) -> torch.Tensor:
        self._fill_state_dict(data.device)
        return self.normalizers[target_name].inverse(data)

    def compute_statistics(self, targets: Dict[str, torch.Tensor]) -> Dict:
        stats = {}
        for name in self.target_names:
            t = targets[name]
            assert t.ndim == 2
            self.normalizers[name].compute_statistics(t)
            stats[name] = {k: v.cpu() for k, v in self.normalizers[name].state_dict().items()}
        self.dataset_statistics_loaded = True
        return stats


class TensorScalarTargetTransform(nn.Module):
    """Wrapper over both (reference transform.py:622-700)."""

    def __init__(self, *, tensor_target_name: Optional[str] = None, tensor_irreps: str = None,
                 scalar_target_names: Optional[List[str]] = None, dataset_statistics_path: Union[str, Path] = None):
        super().__init__()
        self.tensor_normalizer = None if tensor_target_name is None else TensorTargetTransform(
            tensor_target_name, dataset_statistics_path, irreps=tensor_irreps)
        self.scalar_normalizer = None if not scalar_target_names else ScalarTargetTransform(
            scalar_target_names, dataset_statistics_path)

    def forward(self, targets: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        if self.tensor_normalizer is not None:
            name = self.tensor_normalizer.target_name
            targets[name] = self.tensor_normalizer(targets[name])
        if self.scalar_normalizer is not None:
            targets = self.scalar_normalizer(targets)
        return targets

    def inverse(self, data: torch.Tensor, target_name: str) -> torch.Tensor:
        if self.tensor_normalizer is not None and target_name == self.tensor_normalizer.target_name:
            return self.tensor_normalizer.inverse(data)
        if self.scalar_normalizer is not None and target_name in self.scalar_normalizer.target_names:
            return self.scalar_normalizer.inverse(data, target_name)
        raise ValueError(f"Unsupported target name: {target_name}")
