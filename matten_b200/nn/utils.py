"""UVU tensor product, Gate activation and normalisation layers (reference
src/matten/nn/utils.py).  Same constructor arguments, ``irreps_in/irreps_out`` and
``state_dict`` layout as the reference; the arithmetic runs in the CUDA kernels."""
import math
from typing import Callable, Dict, Optional, Union

import torch
from torch import Tensor

from .. import functional as F
from .. import o3, ops
from ..data.irreps import DataKey, ModuleIrreps
from ..o3 import Irreps
from ..plan import GatePlan, UVUPlan, batchnorm_channel_map, linear_blocks, tp_path_exists  # noqa: F401

# name table of reference src/matten/nn/utils.py:14-26 (values are activation names here)
ACTIVATION = {
    "e": {"ssp": "ssp", "silu": "silu", "sigmoid": "sigmoid"},
    "o": {"abs": "abs", "tanh": "tanh"},
}


def _act_name(a: Union[str, Callable]) -> str:
    if isinstance(a, str):
        if a not in o3.ACT_FUNCS:
            raise ValueError(f"unknown activation {a}")
        return a
    for name, (f, _) in o3.ACT_FUNCS.items():
        if a is f:
            return name
    n = getattr(a, "__name__", type(a).__name__).lower()
    for k in ("silu", "tanh", "sigmoid", "abs"):
        if k in n:
            return k
    if "softplus" in n or "ssp" in n:
        return "ssp"
    raise ValueError(f"activation {a} has no CUDA implementation")


class _Buffers(torch.nn.Module):
    """int32/float device tables that must follow ``.to(device)`` but stay out of the state_dict."""

    def _reg(self, name, t):
        self.register_buffer(name, t, persistent=False)


class SpeciesLinear(_Buffers):
    """e3nn ``FullyConnectedTensorProduct(irreps_in, "Sx0e", irreps_out)`` applied to one-hot
    species attributes == species-indexed irreps-wise linear (reference src/matten/nn/conv.py:
    59-61,77-79,84-86).  ``weight`` is the flat e3nn parameter (instruction order, each path
    ``[mul_in, S, mul_out]``)."""

    def __init__(self, irreps_in, num_species: int, irreps_out):
        super().__init__()
        self.irreps_in, self.irreps_out = Irreps(irreps_in), Irreps(irreps_out)
        self.num_species = num_species
        blocks, numel = linear_blocks(self.irreps_in, self.irreps_out, num_species)
        self.handle = ops.LinPlanHandle(blocks, self.irreps_in.dim, self.irreps_out.dim, num_species, numel)
        self.weight_numel = numel
        self.weight = torch.nn.Parameter(torch.randn(numel))

    def forward(self, x: Tensor, species_perm=None, species_ptr=None, residual=None) -> Tensor:
        if self.num_species > 1 and species_ptr is None:
            raise ValueError("SpeciesLinear needs the species grouping of the batch (GraphCache.species_groups)")
        return F.linear(self.handle, x, self.weight, species_perm, species_ptr, residual)


class IrrepsLinear(_Buffers):
    """e3nn ``o3.Linear(irreps_in, irreps_out)`` (no bias) -- reference
    src/matten/nn/nodewise.py:111 and model_factory/tfn_scalar_tensor.py:50."""

    def __init__(self, irreps_in, irreps_out):
        super().__init__()
        self.irreps_in, self.irreps_out = Irreps(irreps_in), Irreps(irreps_out)
        blocks, numel = linear_blocks(self.irreps_in, self.irreps_out, 1)
        self.handle = ops.LinPlanHandle(blocks, self.irreps_in.dim, self.irreps_out.dim, 1, numel)
        self.weight_numel = numel
        self.weight = torch.nn.Parameter(torch.randn(numel))

    def forward(self, x: Tensor) -> Tensor:
        return F.linear(self.handle, x, self.weight)


class _MLPLayer(torch.nn.Module):
    def __init__(self, h_in, h_out):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.randn(h_in, h_out))


class RadialMLP(torch.nn.Module):
    """Parameter container with the layout of ``e3nn.nn.FullyConnectedNet`` (``layer{i}.weight``
    of shape [h_in, h_out]); evaluated inside the fused convolution kernel."""

    def __init__(self, hs, act: str):
        super().__init__()
        self.hs = list(hs)
        self.act_name = act
        f, self.act_id = o3.ACT_FUNCS[act]
        self.act_cst = o3.normalize2mom_const(f)
        for i, (h1, h2) in enumerate(zip(self.hs, self.hs[1:])):
            setattr(self, f"layer{i}", _MLPLayer(h1, h2))

    def weights(self):
        return [getattr(self, f"layer{i}").weight for i in range(len(self.hs) - 1)]


class UVUTensorProduct(torch.nn.Module):
    """reference src/matten/nn/utils.py:170-277.

    ``forward(data1, data2, data_weight)`` keeps the reference's per-edge contract
    ([E, D_in], [E, D_sh], [E, n_rad] -> messages [E, D_mid]); inside ``PointConv`` the same
    kernel is driven with the receiver CSR instead (``fused``), so messages are reduced on chip."""

    def __init__(self, irreps_in1: Irreps, irreps_in2: Irreps, irreps_out: Irreps, *,
                 internal_and_share_weights: bool = False, mlp_input_size: int = None, mlp_hidden_size: int = 8,
                 mlp_num_hidden_layers: int = 1, mlp_activation: Union[str, Callable] = "ssp"):
        super().__init__()
        if internal_and_share_weights:
            raise NotImplementedError("internal shared tensor-product weights are not used by any matten model")
        assert mlp_input_size is not None, ("Expect `mlp_input_size` be provided when "
                                            "`internal_and_share_weights` is set to `False`, got `None`")
        self.plan = UVUPlan(irreps_in1, irreps_in2, irreps_out)
        self.irreps_mid = self.plan.irreps_mid
        sizes = [mlp_input_size] + mlp_num_hidden_layers * [mlp_hidden_size] + [self.plan.weight_numel]
        self.weight_nn = RadialMLP(sizes, _act_name(mlp_activation))
        self._handles = {}

    def handle(self, device) -> ops.ConvPlanHandle:
        key = str(device)
        h = self._handles.get(key)
        if h is None:
            h = ops.ConvPlanHandle(self.plan, self.weight_nn.hs, self.weight_nn.act_id, self.weight_nn.act_cst,
                                   device)
            self._handles[key] = h
        return h

    def fused(self, x: Tensor, sh: Tensor, emb: Tensor, graph, avg_num_neighbors, num_neigh=None) -> Tensor:
        return F.conv(self.handle(x.device), x, sh, emb, self.weight_nn.weights(), graph, avg_num_neighbors,
                      num_neigh)

    def forward(self, data1: Tensor, data2: Tensor, data_weight: Optional[Tensor] = None) -> Tensor:
        assert data_weight is not None, "data for weight not provided"
        E = data1.shape[0]

        class _Identity:  # every edge is its own receiver: no reduction, scale 1
            pass

        g = _Identity()
        g.rowptr = torch.arange(E + 1, dtype=torch.int32, device=data1.device)
        g.perm = torch.arange(E, dtype=torch.int32, device=data1.device)
        g.src_sorted = g.perm
        return F.conv(self.handle(data1.device), data1, data2, data_weight, self.weight_nn.weights(), g, 1.0)

    @property
    def irreps_out(self):
        return self.irreps_mid.simplify()


class _GateModule(_Buffers):
    """e3nn ``Gate`` as element tables (see plan.GatePlan)."""

    def __init__(self, plan: GatePlan):
        super().__init__()
        self.plan = plan
        self.irreps_in, self.irreps_out = plan.irreps_in, plan.irreps_out
        self._reg("src_idx", plan.src_idx)
        self._reg("gate_idx", plan.gate_idx)
        self._reg("act_id", plan.act_id)
        self._reg("act_cst64", plan.act_cst)
        self._reg("inv_first", plan.inv_first)
        self._reg("inv_count", plan.inv_count)
        self._cst = {}

    def tables(self, dtype):
        cst = self._cst.get((dtype, self.act_cst64.device))
        if cst is None:
            cst = self.act_cst64.to(dtype)
            self._cst = {(dtype, self.act_cst64.device): cst}
        return (self.plan.in_dim, self.plan.out_dim, self.src_idx, self.gate_idx, self.act_id, cst, self.inv_first,
                self.inv_count)

    def forward(self, x, affine_a=None, affine_b=None):
        return F.gate(x, self.tables(x.dtype), affine_a, affine_b)


def channel_tables(irreps: Irreps, scalar_test):
    """Per-channel (first column, 2l+1, index among the channels ``scalar_test`` accepts or -1) int32 tables."""
    first, cdim, scal = [], [], []
    col = ns = 0
    for mul, ir in irreps:
        for _ in range(mul):
            first.append(col)
            cdim.append(ir.dim)
            if scalar_test(ir):
                scal.append(ns)
                ns += 1
            else:
                scal.append(-1)
            col += ir.dim
    t = lambda v: torch.tensor(v, dtype=torch.int32)  # noqa: E731
    return t(first), t(cdim), t(scal), ns


class _NormActModule(_Buffers):
    """e3nn ``NormActivation(irreps, f, normalize=True, epsilon=1e-8, bias=False)`` on the scalars + gated irreps
    the tensor product can reach (reference src/matten/nn/utils.py:96-118,142-150): every channel is scaled by
    f(|x|) / |x| with f the raw even-scalar activation."""

    epsilon = 1e-8

    def __init__(self, tp_irreps_in1, tp_irreps_in2, tp_irreps_out, act_name: str):
        super().__init__()
        out = Irreps(tp_irreps_out).sort().irreps.simplify()
        ok = lambda ir: tp_path_exists(tp_irreps_in1, tp_irreps_in2, ir)  # noqa: E731
        scalars = Irreps([(m, ir) for m, ir in out if ir.l == 0 and ok(ir)])
        gated = Irreps([(m, ir) for m, ir in out if ir.l > 0 and ok(ir)])
        self.irreps_in = self.irreps_out = (scalars + gated).simplify()
        self.act_id = o3.ACT_FUNCS[act_name][1]
        first, cdim, _, _ = channel_tables(self.irreps_in, lambda ir: False)
        self._reg("chan_first", first)
        self._reg("chan_dim", cdim)

    def forward(self, x, affine_a=None, affine_b=None):
        tables = (self.chan_first, self.chan_dim)
        if torch.is_grad_enabled() and x.requires_grad:
            from .. import autograd as A

            y = A.NormActFn.apply(x, tables, self.act_id, self.epsilon)
        else:
            y = ops.norm_act(x.detach(), tables, self.act_id, self.epsilon)
        if affine_a is not None:
            y = A_affine(y, affine_a, affine_b)
        return y


class ActivationLayer(torch.nn.Module):
    """reference src/matten/nn/utils.py:29-167 (``gate`` and ``norm``)."""

    def __init__(self, tp_irreps_in1: Irreps, tp_irreps_in2: Irreps, tp_irreps_out: Irreps, *,
                 activation_type: str = "gate", activation_scalars: Dict[str, str] = None,
                 activation_gates: Dict[str, str] = None):
        super().__init__()
        km = {"e": 1, "o": -1}
        if activation_scalars is None:
            a_s = {1: "ssp", -1: "tanh"}
        else:
            a_s = {km[k]: _act_name(ACTIVATION[k][v]) for k, v in activation_scalars.items()}
        if activation_gates is None:
            a_g = {1: "ssp", -1: "abs"}
        else:
            a_g = {km[k]: _act_name(ACTIVATION[k][v]) for k, v in activation_gates.items()}
        if activation_type == "gate":
            self.activation = _GateModule(GatePlan(tp_irreps_in1, tp_irreps_in2, tp_irreps_out, a_s, a_g))
        elif activation_type == "norm":
            self.activation = _NormActModule(tp_irreps_in1, tp_irreps_in2, tp_irreps_out, a_s[1])
        else:
            raise ValueError(f"Support `activation_type` includes ('gate', 'norm'), got {activation_type}")

    def forward(self, x: Tensor, affine_a=None, affine_b=None) -> Tensor:
        return self.activation(x, affine_a, affine_b)

    @property
    def irreps_in(self):
        return self.activation.irreps_in

    @property
    def irreps_out(self):
        return self.activation.irreps_out


class BatchNorm(_Buffers):
    """e3nn ``nn.BatchNorm(irreps)``: eps 1e-5, momentum 0.1, affine, component normalisation,
    mean reduce (reference src/matten/nn/utils.py:418).  Only 0e channels are centred/biased."""

    def __init__(self, irreps, eps=1e-5, momentum=0.1):
        super().__init__()
        self.irreps = Irreps(irreps)
        self.eps, self.momentum = eps, momentum
        feat, scal, nf, ns = batchnorm_channel_map(self.irreps)
        self.num_features, self.num_scalar = nf, ns
        self._reg("feat_idx", feat)
        self._reg("scal_idx", scal.clamp(min=0))
        self._reg("scal_mask", (scal >= 0))
        self._reg("scal_cols", torch.nonzero(scal >= 0).reshape(-1))
        self.register_buffer("running_mean", torch.zeros(ns))
        self.register_buffer("running_var", torch.ones(nf))
        self.weight = torch.nn.Parameter(torch.ones(nf))
        self.bias = torch.nn.Parameter(torch.zeros(ns))

    def channel_matrix(self, dtype):
        """[D, num_features] 0/1 matrix summing the 2l+1 columns of every channel (deterministic, unlike index_add)."""
        key = (dtype, self.feat_idx.device)
        if getattr(self, "_chan_key", None) != key:
            D = self.feat_idx.shape[0]
            m = torch.zeros((D, self.num_features), dtype=dtype, device=self.feat_idx.device)
            m[torch.arange(D, device=m.device), self.feat_idx] = 1
            self._chan, self._chan_key = m, key
        return self._chan

    def eval_affine(self, dtype):
        """Per-element (a, b) with ``y = a*x + b`` for eval mode: tiny [D] vectors derived from
        the running statistics (parameter plumbing, not per-node work).  Cached between calls while the parameters
        and running statistics are unchanged (tensor version counters) -- ten small launches per layer otherwise."""
        key = (dtype, self.feat_idx.device, self.weight._version, self.bias._version, self.running_mean._version,
               self.running_var._version, self.weight.data_ptr(), self.running_var.data_ptr())
        cached = getattr(self, "_affine_cache", None)
        if cached is not None and cached[0] == key and not (torch.is_grad_enabled() and self.weight.requires_grad):
            return cached[1]
        rstd = (self.running_var + self.eps).pow(-0.5) * self.weight
        a = rstd[self.feat_idx]
        if self.num_scalar > 0:
            shift = self.bias - self.running_mean * rstd[self._scalar_channels()]
            b = torch.where(self.scal_mask, shift[self.scal_idx], torch.zeros_like(a))
        else:
            b = torch.zeros_like(a)
        out = (a.to(dtype).contiguous(), b.to(dtype).contiguous())
        if not (torch.is_grad_enabled() and self.weight.requires_grad):
            self._affine_cache = (key, (out[0].detach(), out[1].detach()))
        return out

    def _scalar_channels(self):
        if not hasattr(self, "_sc_cache") or self._sc_cache.device != self.feat_idx.device:
            ch = []
            f = 0
            for m, ir in self.irreps:
                if ir.is_scalar():
                    ch += list(range(f, f + m))
                f += m
            self._sc_cache = torch.tensor(ch, dtype=torch.int64, device=self.feat_idx.device)
        return self._sc_cache

    def forward(self, x: Tensor) -> Tensor:
        if self.training:
            from .. import autograd as A

            return A.batchnorm_train(self, x)
        a, b = self.eval_affine(x.dtype)
        # identity gate tables are not needed: y = a*x + b through the gate kernel's affine stage
        return A_affine(x, a, b)


def A_affine(x, a, b):
    if torch.is_grad_enabled() and (x.requires_grad or a.requires_grad or b.requires_grad):
        from .. import autograd as A

        return A.AffineFn.apply(x, a, b)
    D = x.shape[-1]
    key = (D, x.device)
    t = _IDENT.get(key)
    if t is None:
        src = torch.arange(D, dtype=torch.int32, device=x.device)
        gate = torch.full((D,), -1, dtype=torch.int32, device=x.device)
        act = torch.zeros(D, dtype=torch.int32, device=x.device)
        t = (src, gate, act)
        _IDENT[key] = t
    cst = torch.ones(D, dtype=x.dtype, device=x.device)
    return F.gate(x, (D, D, t[0], t[1], t[2], cst, None, None), a, b)


_IDENT = {}


class InstanceNorm(_Buffers):
    """Graph-wise instance normalisation (reference src/matten/nn/utils.py:448-588): every graph is an instance and
    its nodes are the samples.  l = 0 channels of either parity are centred and biased (the reference tests
    ``ir.l == 0``); the same graph statistics are used in training and evaluation (reference note at :440-441).

    ``forward(input, batch)`` takes the sorted node -> graph vector like the reference; ``graph_ptr`` (the CSR form
    the kernels consume) can be passed instead when the caller already holds it."""

    def __init__(self, irreps, eps=1e-5, affine=True, reduce="mean", normalization="component"):
        super().__init__()
        self.irreps = Irreps(irreps)
        self.eps, self.affine = eps, affine
        assert isinstance(reduce, str), "reduce should be passed as a string value"
        assert reduce in ["mean", "max"], "reduce needs to be 'mean' or 'max'"
        assert normalization in ["norm", "component"], "normalization needs to be 'norm' or 'component'"
        self.reduce, self.normalization = reduce, normalization
        first, cdim, scal, ns = channel_tables(self.irreps, lambda ir: ir.l == 0)
        self.num_features, self.num_scalar = first.shape[0], ns
        self._reg("chan_first", first)
        self._reg("chan_dim", cdim)
        self._reg("chan_scalar", scal)
        self._reg("scalar_channels", torch.nonzero(scal >= 0).reshape(-1))
        if affine:
            self.weight = torch.nn.Parameter(torch.ones(self.num_features))
            self.bias = torch.nn.Parameter(torch.zeros(self.num_scalar))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)

    def __repr__(self):
        return f"{self.__class__.__name__} ({self.irreps}, eps={self.eps})"

    def tables(self):
        return self.chan_first, self.chan_dim, self.chan_scalar

    def forward(self, input: Tensor, batch: Tensor = None, graph_ptr: Tensor = None) -> Tensor:
        if input.shape[-1] != self.irreps.dim:
            raise AssertionError(f"`ix` should have reached input.size(-1) ({input.shape[-1]}), "
                                 f"but it ended at {self.irreps.dim}")
        if graph_ptr is None:
            if batch is None:
                raise ValueError("InstanceNorm needs the node -> graph vector `batch`")
            flag = ops.new_flag(input.device)
            ops.check_sorted(batch, flag)
            num_graphs = int(batch[-1].item()) + 1 if batch.numel() else 0
            graph_ptr, _ = ops.csr_by_key(batch, num_graphs, False, flag)
            ops.raise_on_flag(flag)
        w = self.weight if self.affine else None
        b = self.bias if self.affine else None
        grad = torch.is_grad_enabled() and (input.requires_grad or (self.affine and self.weight.requires_grad))
        if grad:
            from .. import autograd as A

            return A.InstanceNormFn.apply(input, w, b, graph_ptr, self)
        return ops.instance_norm_fwd(input.detach(), graph_ptr, self.tables(), None if w is None else w.detach(),
                                     None if b is None else b.detach(), self.eps, self.reduce,
                                     self.normalization)[0]


class NormalizationLayer(torch.nn.Module):
    """reference src/matten/nn/utils.py:397-437"""

    def __init__(self, irreps: Irreps, method: str = None):
        super().__init__()
        self.method = method
        supported = ("batch", "instance", "none", None)
        assert method in supported, f"Unsupported normalization {method}"
        if method == "batch":
            self.n = BatchNorm(irreps)
        elif method == "instance":
            self.n = InstanceNorm(irreps)
        else:
            self.n = None

    def forward(self, x: Tensor, batch: Tensor = None, graph_ptr: Tensor = None) -> Tensor:
        if self.method == "batch":
            x = self.n(x)
        elif self.method == "instance":
            x = self.n(x, batch, graph_ptr)
        return x
