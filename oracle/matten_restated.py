"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Not imported by the product package.

CPU restatement of the matten hot path (reference src/matten/nn/*.py,
src/matten/model_factory/*.py) on top of ``oracle.e3nn_restated``.  Module
and parameter names follow the reference so that one ``state_dict`` drives both
this oracle and the CUDA modules under test.  PARITY UNPINNED against e3nn
(see the header of e3nn_restated.py); pinned against the reference's own tests
(index symmetry + rotation equivariance, tests/model/test_tfn_tensor.py:98-139;
atomic-number KAT, tests/nn/test_embedding.py:6-13).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference
arm may import this file.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional

import torch

from . import e3nn_restated as E

# data-dict keys: reference src/matten/data/_key.py:14-49
POSITIONS = "pos"
NODE_ATTRS = "node_attrs"
NODE_FEATURES = "node_features"
EDGE_INDEX = "edge_index"
EDGE_CELL_SHIFT = "edge_cell_shift"
EDGE_VECTORS = "edge_vectors"
EDGE_LENGTH = "edge_lengths"
EDGE_ATTRS = "edge_attrs"
EDGE_EMBEDDING = "edge_embedding"
CELL = "cell"
NUM_NEIGH = "num_neigh"
ATOMIC_NUMBERS = "atomic_numbers"
SPECIES_INDEX = "species_index"
BATCH = "batch"
OUT_FIELD_NAME = "my_model_output"  # model_factory/tfn_scalar_tensor.py:29


def _ssp(x):
    return torch.nn.functional.softplus(x) - math.log(2.0)


# reference src/matten/nn/utils.py:14-26
ACTIVATION = {
    "e": {"ssp": _ssp, "silu": torch.nn.functional.silu, "sigmoid": torch.sigmoid},
    "o": {"abs": torch.abs, "tanh": torch.tanh},
}


def tp_path_exists(irreps_in1, irreps_in2, ir_out) -> bool:
    """reference src/matten/nn/utils.py:358-367"""
    a = E.irreps_simplify(E.parse_irreps(irreps_in1))
    b = E.irreps_simplify(E.parse_irreps(irreps_in2))
    if isinstance(ir_out, str):
        ((_, lo, po),) = E.parse_irreps(ir_out)
    else:
        lo, po = ir_out
    return any((lo, po) in E.irrep_product(l1, p1, l2, p2) for _, l1, p1 in a for _, l2, p2 in b)


# ------------------------------------------------------------- edge geometry --
def with_edge_vectors(data: Dict[str, torch.Tensor], with_lengths=True):
    """reference src/matten/nn/_nequip.py:214-268"""
    if EDGE_VECTORS not in data:
        pos, ei = data[POSITIONS], data[EDGE_INDEX]
        vec = pos[ei[1]] - pos[ei[0]]
        if CELL in data:
            cell = data[CELL].view(-1, 3, 3)
            shift = data[EDGE_CELL_SHIFT]
            if cell.shape[0] > 1:
                vec = vec + torch.einsum("ni,nij->nj", shift, cell[data[BATCH][ei[0]]])
            else:
                vec = vec + torch.einsum("ni,ij->nj", shift, cell.squeeze(0))
        data[EDGE_VECTORS] = vec
    if with_lengths and EDGE_LENGTH not in data:
        data[EDGE_LENGTH] = torch.linalg.norm(data[EDGE_VECTORS], dim=-1)
    return data


class SphericalHarmonicEdgeAttrs(torch.nn.Module):
    """reference src/matten/nn/_nequip.py:130-176"""

    def __init__(self, irreps_edge_sh, irreps_in=None):
        super().__init__()
        sh = E.parse_irreps(irreps_edge_sh) if isinstance(irreps_edge_sh, str) else \
            [(1, l, (-1) ** l) for l in range(irreps_edge_sh + 1)]
        assert sh == [(1, l, (-1) ** l) for l in range(len(sh))], "oracle supports full sh only"
        self.lmax = len(sh) - 1
        self.irreps_out = dict(irreps_in or {})
        self.irreps_out[EDGE_ATTRS] = sh

    def forward(self, data):
        data = with_edge_vectors(data, with_lengths=False)
        data[EDGE_ATTRS] = E.spherical_harmonics(self.lmax, data[EDGE_VECTORS], True, "component")
        return data


class EdgeLengthEmbedding(torch.nn.Module):
    """reference src/matten/nn/embedding.py:158-203"""

    def __init__(self, irreps_in=None, num_basis=10, start=0.0, end=5.0, basis="bessel", cutoff=True):
        super().__init__()
        assert basis == "bessel"
        self.num_basis, self.start, self.end, self.cutoff = num_basis, start, end, cutoff
        self.irreps_out = dict(irreps_in or {})
        self.irreps_out[EDGE_EMBEDDING] = [(num_basis, 0, 1)]

    def forward(self, data):
        data = with_edge_vectors(data, with_lengths=True)
        emb = E.soft_one_hot_linspace_bessel(data[EDGE_LENGTH], self.start, self.end,
                                             self.num_basis, self.cutoff)
        data[EDGE_EMBEDDING] = emb.mul(self.num_basis**0.5)
        return data


class PolynomialCutoff(torch.nn.Module):
    """reference src/matten/nn/_nequip.py:43-76"""

    def __init__(self, r_max, p=6):
        super().__init__()
        self.p, self.r_max = float(p), float(r_max)

    def forward(self, x):
        p, u = self.p, x / self.r_max
        env = (1.0 - ((p + 1.0) * (p + 2.0) / 2.0) * torch.pow(u, p)
               + p * (p + 2.0) * torch.pow(u, p + 1.0)
               - (p * (p + 1.0) / 2) * torch.pow(u, p + 2.0))
        return env * (x < self.r_max).to(x.dtype)


class BesselBasis(torch.nn.Module):
    """reference src/matten/nn/_nequip.py:80-126"""

    def __init__(self, r_max, num_basis=8, trainable=True):
        super().__init__()
        self.r_max, self.num_basis = float(r_max), num_basis
        w = torch.linspace(1.0, num_basis, num_basis) * math.pi
        if trainable:
            self.bessel_weights = torch.nn.Parameter(w)
        else:
            self.register_buffer("bessel_weights", w)

    def forward(self, x):
        num = torch.sin(self.bessel_weights * x.unsqueeze(-1) / self.r_max)
        return (2.0 / self.r_max) * (num / x.unsqueeze(-1))


class RadialBasisEdgeEncoding(torch.nn.Module):
    """reference src/matten/nn/_nequip.py:180-210"""

    def __init__(self, basis_kwargs, cutoff_kwargs, irreps_in=None):
        super().__init__()
        self.basis = BesselBasis(**basis_kwargs)
        self.cutoff = PolynomialCutoff(**cutoff_kwargs)
        self.irreps_out = dict(irreps_in or {})
        self.irreps_out[EDGE_EMBEDDING] = [(self.basis.num_basis, 0, 1)]

    def forward(self, data):
        data = with_edge_vectors(data, with_lengths=True)
        r = data[EDGE_LENGTH]
        data[EDGE_EMBEDDING] = self.basis(r) * self.cutoff(r)[:, None]
        return data


# ---------------------------------------------------------- species embedding --
class _AtomicNumberToIndex(torch.nn.Module):
    """reference src/matten/nn/embedding.py:206-263"""

    def __init__(self, allowed_atomic_numbers: List[int]):
        super().__init__()
        allowed = torch.as_tensor(sorted(allowed_atomic_numbers), dtype=torch.long)
        n = len(allowed)
        self.register_buffer("_min_Z", allowed.min())
        self.register_buffer("_max_Z", allowed.max())
        self.register_buffer("_num_species", torch.as_tensor(n))
        lut = torch.full((1 + int(self._max_Z) - int(self._min_Z),), -1, dtype=torch.long)
        lut[allowed - self._min_Z] = torch.arange(n)
        self.register_buffer("_Z_to_index", lut)

    def forward(self, z):
        if z.min() < self._min_Z or z.max() > self._max_Z:
            raise RuntimeError("Invalid atomic numbers.")
        idx = self._Z_to_index[z - self._min_Z]
        if idx.min() < 0:
            raise RuntimeError("Invalid atomic numbers.")
        return idx

    @property
    def num_species(self):
        return int(self._num_species)


class SpeciesEmbedding(torch.nn.Module):
    """reference src/matten/nn/embedding.py:12-110"""

    def __init__(self, irreps_in=None, embedding_dim=16, allowed_species=None, **_):
        super().__init__()
        self.atomic_number_to_index = _AtomicNumberToIndex(allowed_species)
        self.num_species = self.atomic_number_to_index.num_species
        self.linear = torch.nn.Linear(self.num_species, embedding_dim)
        self.irreps_out = dict(irreps_in or {})
        self.irreps_out[NODE_ATTRS] = [(self.num_species, 0, 1)]
        self.irreps_out[NODE_FEATURES] = [(embedding_dim, 0, 1)]

    def forward(self, data):
        if SPECIES_INDEX in data:
            t = data[SPECIES_INDEX]
        else:
            t = self.atomic_number_to_index(data[ATOMIC_NUMBERS])
            data[SPECIES_INDEX] = t
        attrs = torch.nn.functional.one_hot(t, num_classes=self.num_species).to(self.linear.weight.dtype)
        data[NODE_ATTRS] = attrs
        data[NODE_FEATURES] = self.linear(attrs)
        return data


# --------------------------------------------------------------- convolution --
class UVUTensorProduct(torch.nn.Module):
    """reference src/matten/nn/utils.py:170-277"""

    def __init__(self, irreps_in1, irreps_in2, irreps_out, *, mlp_input_size, mlp_hidden_size=8,
                 mlp_num_hidden_layers=1, mlp_activation=ACTIVATION["e"]["ssp"]):
        super().__init__()
        in1, in2, out = map(E.parse_irreps, (irreps_in1, irreps_in2, irreps_out))
        mid, instr = [], []
        out_set = {(l, p) for _, l, p in out}
        for i, (mul, l1, p1) in enumerate(in1):
            for j, (_, l2, p2) in enumerate(in2):
                for lo, po in E.irrep_product(l1, p1, l2, p2):
                    if (lo, po) in out_set:  # the `== Irreps("0e")` clause is always False
                        instr.append((i, j, len(mid), "uvu", True))
                        mid.append((mul, lo, po))
        assert E.irreps_dim(mid) > 0
        self.irreps_mid, perm = E.irreps_sort(mid)
        instr = [(a, b, perm[c], m, t) for a, b, c, m, t in instr]
        self.tp = E.TensorProduct(in1, in2, self.irreps_mid, instr,
                                  internal_weights=False, shared_weights=False)
        sizes = [mlp_input_size] + mlp_num_hidden_layers * [mlp_hidden_size] + [self.tp.weight_numel]
        self.weight_nn = E.FullyConnectedNet(sizes, act=mlp_activation)

    def forward(self, data1, data2, data_weight):
        return self.tp(data1, data2, self.weight_nn(data_weight))

    @property
    def irreps_out(self):
        return E.irreps_simplify(self.irreps_mid)


class PointConv(torch.nn.Module):
    """reference src/matten/nn/conv.py:26-143"""

    def __init__(self, irreps_in, conv_layer_irreps, fc_num_hidden_layers=1, fc_hidden_size=8,
                 avg_num_neighbors=None):
        super().__init__()
        self.avg_num_neighbors = avg_num_neighbors
        x_ir, a_ir, sh_ir = irreps_in[NODE_FEATURES], irreps_in[NODE_ATTRS], irreps_in[EDGE_ATTRS]
        conv_ir = E.parse_irreps(conv_layer_irreps)
        self.lin1 = E.FullyConnectedTensorProduct(x_ir, a_ir, x_ir)
        self.tp = UVUTensorProduct(x_ir, sh_ir, conv_ir,
                                   mlp_input_size=E.irreps_dim(irreps_in[EDGE_EMBEDDING]),
                                   mlp_hidden_size=fc_hidden_size,
                                   mlp_num_hidden_layers=fc_num_hidden_layers,
                                   mlp_activation=ACTIVATION["e"]["silu"])
        self.lin2 = E.FullyConnectedTensorProduct(self.tp.irreps_out, a_ir, conv_ir)
        self.sc = E.FullyConnectedTensorProduct(x_ir, a_ir, conv_ir)
        self.irreps_out = dict(irreps_in)
        self.irreps_out[NODE_FEATURES] = conv_ir

    def forward(self, data):
        x, attrs = data[NODE_FEATURES], data[NODE_ATTRS]
        src, dst = data[EDGE_INDEX]
        sc = self.sc(x, attrs)
        x = self.lin1(x, attrs)
        msg = self.tp(x[src], data[EDGE_ATTRS], data[EDGE_EMBEDDING])
        agg = E.scatter(msg, dst, dim_size=len(x))
        if self.avg_num_neighbors is not None:
            agg = agg.div(self.avg_num_neighbors**0.5)
        else:
            agg = agg.div(data[NUM_NEIGH].reshape(-1, 1) ** 0.5)
        data[NODE_FEATURES] = sc + self.lin2(agg, attrs)
        return data


class ActivationLayer(torch.nn.Module):
    """reference src/matten/nn/utils.py:29-167"""

    def __init__(self, tp_irreps_in1, tp_irreps_in2, tp_irreps_out, *, activation_type="gate",
                 activation_scalars=None, activation_gates=None):
        super().__init__()
        km = {"e": 1, "o": -1}
        if activation_scalars is None:
            a_s = {1: ACTIVATION["e"]["ssp"], -1: ACTIVATION["o"]["tanh"]}
        else:
            a_s = {km[k]: ACTIVATION[k][v] for k, v in activation_scalars.items()}
        if activation_gates is None:
            a_g = {1: ACTIVATION["e"]["ssp"], -1: ACTIVATION["o"]["abs"]}
        else:
            a_g = {km[k]: ACTIVATION[k][v] for k, v in activation_gates.items()}
        out, _ = E.irreps_sort(E.parse_irreps(tp_irreps_out))
        out = E.irreps_simplify(out)
        ok = lambda l, p: tp_path_exists(tp_irreps_in1, tp_irreps_in2, (l, p))  # noqa: E731
        scalars = [(m, l, p) for m, l, p in out if l == 0 and ok(l, p)]
        gated = [(m, l, p) for m, l, p in out if l > 0 and ok(l, p)]
        if activation_type == "norm":
            # norm is an even scalar, so activation_scalars[1]   (utils.py:142-150)
            self.activation = E.NormActivation(E.irreps_simplify(scalars + gated), a_s[1], normalize=True,
                                               epsilon=1e-8, bias=False)
            return
        if activation_type != "gate":
            raise ValueError(f"Support `activation_type` includes ('gate', 'norm'), got {activation_type}")
        if E.irreps_dim(gated) > 0:
            if ok(0, 1):
                gp = 1
            elif ok(0, -1):
                gp = -1
            else:
                raise ValueError("unable to produce gates")
            gates = E.irreps_simplify([(m, 0, gp) for m, _, _ in gated])
        else:
            gates = []
        self.activation = E.Gate(scalars, [a_s[p] for _, _, p in scalars],
                                 gates, [a_g[p] for _, _, p in gates], gated)

    def forward(self, x):
        return self.activation(x)

    @property
    def irreps_in(self):
        return self.activation.irreps_in

    @property
    def irreps_out(self):
        return self.activation.irreps_out


def _graph_pool(x, batch, reduce):
    """torch_geometric global_mean_pool / global_max_pool: scatter over the node -> graph vector."""
    G = int(batch.max()) + 1 if batch.numel() else 0
    return E.scatter(x, batch, dim_size=G, reduce=reduce)


class InstanceNorm(torch.nn.Module):
    """reference src/matten/nn/utils.py:448-588: each graph is an instance, its nodes the samples; l == 0 channels
    (either parity) are centred by the graph mean and get a bias; the channel's squared norm (component mean or
    sum) is pooled over the graph's nodes by mean or max; the same statistics serve training and evaluation."""

    def __init__(self, irreps, eps=1e-5, affine=True, reduce="mean", normalization="component"):
        super().__init__()
        self.irreps = E.parse_irreps(irreps)
        self.eps, self.affine = eps, affine
        ns = sum(m for m, l, p in self.irreps if l == 0)
        nf = sum(m for m, _, _ in self.irreps)
        if affine:
            self.weight = torch.nn.Parameter(torch.ones(nf))
            self.bias = torch.nn.Parameter(torch.zeros(ns))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)
        assert reduce in ["mean", "max"] and normalization in ["norm", "component"]
        self.reduce, self.normalization = reduce, normalization

    def forward(self, input, batch):
        fields = []
        ix = iw = ib = 0
        for m, l, p in self.irreps:
            d = 2 * l + 1
            field = input[:, ix:ix + m * d].reshape(-1, m, d)
            ix += m * d
            if l == 0:
                mean = _graph_pool(field, batch, "mean").reshape(-1, m, 1)
                field = field - mean[batch]
            fn = field.pow(2).sum(-1) if self.normalization == "norm" else field.pow(2).mean(-1)
            fn = _graph_pool(fn, batch, self.reduce)
            fn = (fn + self.eps).pow(-0.5)
            if self.affine:
                fn = fn * self.weight[None, iw:iw + m]
                iw += m
            field = field * fn[batch].reshape(-1, m, 1)
            if self.affine and d == 1:
                field = field + self.bias[ib:ib + m].reshape(m, 1)
                ib += m
            fields.append(field.reshape(-1, m * d))
        assert ix == input.shape[-1]
        return torch.cat(fields, -1)


class NormalizationLayer(torch.nn.Module):
    """reference src/matten/nn/utils.py:397-437"""

    def __init__(self, irreps, method=None):
        super().__init__()
        assert method in ("batch", "instance", "none", None), f"Unsupported normalization {method}"
        self.method = method
        self.n = E.BatchNorm(irreps) if method == "batch" else InstanceNorm(irreps) if method == "instance" else None

    def forward(self, x, batch):
        if self.method == "batch":
            return self.n(x)
        if self.method == "instance":
            return self.n(x, batch)
        return x


class MeanNormNormalize(torch.nn.Module):
    """reference src/matten/data/transform.py:59-216 (the usable ``reduce="mean"`` branch; ``max`` raises there)."""

    def __init__(self, irreps, mean=None, norm=None, normalization="component", reduce="mean", eps=1e-5, scale=1.0):
        super().__init__()
        self.irreps = E.parse_irreps(irreps)
        self.normalization, self.reduce, self.eps, self.scale = normalization, reduce, eps, scale
        self.mean, self.norm = mean, norm

    def forward(self, data):
        if self.mean is None or self.norm is None:
            raise RuntimeError("mean and norm not initialized.")
        return (data - self.mean) / (self.norm * self.scale)

    def inverse(self, data):
        if self.mean is None or self.norm is None:
            raise RuntimeError("mean and norm not initialized.")
        return data * (self.norm * self.scale) + self.mean

    def compute_statistics(self, data):
        all_mean, all_norm = [], []
        ix = 0
        for m, l, p in self.irreps:
            d = 2 * l + 1
            field = data[:, ix:ix + m * d].reshape(-1, m, d)
            ix += m * d
            if l == 0 and p == 1:
                fm = field.mean(0).reshape(m)
                field = field - fm.reshape(-1, m, 1)
            else:
                fm = torch.zeros(m, dtype=data.dtype)
            all_mean.append(torch.repeat_interleave(fm, d))
            fn = field.pow(2).sum(-1) if self.normalization == "norm" else field.pow(2).mean(-1)
            assert self.reduce == "mean"
            fn = (fn.mean(0) + self.eps).pow(0.5)
            all_norm.append(torch.repeat_interleave(fn, d))
        assert ix == data.shape[-1]
        self.mean, self.norm = torch.cat(all_mean), torch.cat(all_norm)
        return self.mean, self.norm


class ScalarNormalize(MeanNormNormalize):
    """reference src/matten/data/transform.py:219-302; sklearn StandardScaler.fit = column mean and population
    standard deviation, deviations below 10 eps replaced by 1 (sklearn ``_handle_zeros_in_scale``)."""

    def __init__(self, num_features, mean=None, norm=None, scale=1.0):
        torch.nn.Module.__init__(self)
        self.scale, self.mean, self.norm = scale, mean, norm

    def compute_statistics(self, data):
        assert data.ndim == 2, "Can only deal with tensor [N_samples, N_features]"
        x = data.double()
        mean = x.mean(0)
        std = ((x - mean) ** 2).mean(0).sqrt()
        std = torch.where(std < 10 * torch.finfo(torch.float64).eps, torch.ones_like(std), std)
        self.mean, self.norm = mean.to(data.dtype), std.to(data.dtype)
        return self.mean, self.norm


class PointConvWithActivation(torch.nn.Module):
    """reference src/matten/nn/conv.py:146-215"""

    def __init__(self, irreps_in, conv_layer_irreps, fc_num_hidden_layers=1, fc_hidden_size=8,
                 avg_num_neighbors=None, activation_type="gate",
                 activation_scalars={"e": "silu", "o": "tanh"},
                 activation_gates={"e": "sigmoid", "o": "tanh"}, normalization=None):
        super().__init__()
        self.act = ActivationLayer(irreps_in[NODE_FEATURES], irreps_in[EDGE_ATTRS], conv_layer_irreps,
                                   activation_type=activation_type,
                                   activation_scalars=activation_scalars,
                                   activation_gates=activation_gates)
        self.conv = PointConv(irreps_in, self.act.irreps_in, fc_num_hidden_layers, fc_hidden_size,
                              avg_num_neighbors)
        self.norm = NormalizationLayer(self.act.irreps_out, method=normalization)
        self.irreps_out = dict(irreps_in)
        self.irreps_out[NODE_FEATURES] = self.act.irreps_out

    def forward(self, data):
        data = self.conv(data)
        x = self.act(data[NODE_FEATURES])
        data[NODE_FEATURES] = self.norm(x, data[BATCH])
        return data


# --------------------------------------------------------------------- heads --
class NodewiseLinear(torch.nn.Module):
    """reference src/matten/nn/nodewise.py:89-117"""

    def __init__(self, irreps_in, irreps_out=None, field=NODE_FEATURES, out_field=None):
        super().__init__()
        self.field, self.out_field = field, out_field or field
        io = E.parse_irreps(irreps_out) if irreps_out is not None else irreps_in[field]
        self.linear = E.Linear(irreps_in[field], io)
        self.irreps_out = dict(irreps_in)
        self.irreps_out[self.out_field] = io

    def forward(self, data):
        data[self.out_field] = self.linear(data[self.field])
        return data


class NodewiseReduce(torch.nn.Module):
    """reference src/matten/nn/nodewise.py:120-148"""

    def __init__(self, irreps_in, field, out_field=None, reduce="sum"):
        super().__init__()
        self.field, self.reduce = field, reduce
        self.out_field = f"{reduce}_{field}" if out_field is None else out_field
        self.irreps_out = dict(irreps_in)
        self.irreps_out[self.out_field] = irreps_in[field]

    def forward(self, data):
        if BATCH not in data:
            data[BATCH] = torch.zeros(len(data[POSITIONS]), dtype=torch.long)
        data[self.out_field] = E.scatter(data[self.field], data[BATCH], reduce=self.reduce)
        return data


class NodewiseSelect(torch.nn.Module):
    """reference src/matten/nn/nodewise.py:18-86"""

    def __init__(self, irreps_in, field=NODE_FEATURES, out_field=None, mask_field=None):
        super().__init__()
        self.field, self.out_field, self.mask_field = field, out_field or field, mask_field
        self.irreps_out = dict(irreps_in)
        self.irreps_out[self.out_field] = irreps_in[field]

    def forward(self, data):
        data = data.copy()
        v = data[self.field]
        data[self.out_field] = v if self.mask_field is None else v[data[self.mask_field]]
        return data


# ---------------------------------------------------------- model factories --
def create_model(hparams, dataset_hparams, atomic: bool = False) -> torch.nn.Sequential:
    """reference src/matten/model_factory/tfn_scalar_tensor.py:103-195 and
    tfn_atomic_tensor.py:103-199 (atomic=True: no pooling, hidden = CartesianTensor)."""
    layers = OrderedDict()
    m = SpeciesEmbedding(None, hparams["species_embedding_dim"], dataset_hparams["allowed_species"])
    layers["one_hot"] = m
    m = SphericalHarmonicEdgeAttrs(hparams["irreps_edge_sh"], m.irreps_out)
    layers["spharm_edges"] = m
    m = EdgeLengthEmbedding(m.irreps_out, hparams["num_radial_basis"], hparams["radial_basis_start"],
                            hparams["radial_basis_end"], hparams["radial_basis_type"])
    layers["radial_basis"] = m
    nn_ = hparams["average_num_neighbors"]
    if isinstance(nn_, str) and nn_.lower() == "auto":
        nn_ = dataset_hparams["average_num_neighbors"]
    for i in range(hparams["num_layers"]):
        m = PointConvWithActivation(m.irreps_out, hparams["conv_layer_irreps"],
                                    hparams["invariant_layers"], hparams["invariant_neurons"], nn_,
                                    activation_type=hparams["nonlinearity_type"],
                                    normalization=hparams["normalization"])
        layers[f"layer{i}_convnet"] = m
    m = PointConv(m.irreps_out, hparams["conv_layer_irreps"], hparams["invariant_layers"],
                  hparams["invariant_neurons"], nn_)
    layers["conv_layer_last"] = m
    if atomic:
        formula = hparams["output_formula"].lower()
        io = [(1, 0, 1)] if formula == "scalar" else E.CartesianTensor(formula).irreps
        m = NodewiseLinear(m.irreps_out, io, out_field=OUT_FIELD_NAME)
        layers["conv_to_output_hidden"] = m
    else:
        m = NodewiseLinear(m.irreps_out, hparams["conv_to_output_hidden_irreps_out"],
                           out_field=OUT_FIELD_NAME)
        layers["conv_to_output_hidden"] = m
        m = NodewiseReduce(m.irreps_out, OUT_FIELD_NAME, OUT_FIELD_NAME, hparams["reduce"])
        layers["output_pooling"] = m
    return torch.nn.Sequential(layers)


class ScalarTensorModel(torch.nn.Module):
    """backbone + out_layer + optional to_cartesian: the arithmetic of reference
    ScalarTensorModel.decode (model_factory/tfn_scalar_tensor.py:40-79) without the
    Lightning shell."""

    def __init__(self, hparams, dataset_hparams):
        super().__init__()
        self.backbone = create_model(hparams, dataset_hparams)
        formula = hparams["output_formula"].lower()
        self.ct = None if formula == "scalar" else E.CartesianTensor(formula)
        io = [(1, 0, 1)] if formula == "scalar" else self.ct.irreps
        self.extra_layers_dict = torch.nn.ModuleDict(
            {"out_layer": E.Linear(hparams["conv_to_output_hidden_irreps_out"], io)})
        self.cartesian = hparams.get("output_format", "irreps") == "cartesian" and self.ct is not None

    def forward(self, data):
        out = self.backbone(dict(data))[OUT_FIELD_NAME]
        out = self.extra_layers_dict["out_layer"](out)
        return self.ct.to_cartesian(out) if self.cartesian else out


class AtomicTensorModel(torch.nn.Module):
    """reference model_factory/tfn_atomic_tensor.py:31-100 (no pooling, no out_layer)."""

    def __init__(self, hparams, dataset_hparams):
        super().__init__()
        self.backbone = create_model(hparams, dataset_hparams, atomic=True)
        formula = hparams["output_formula"].lower()
        self.ct = None if formula == "scalar" else E.CartesianTensor(formula)
        self.cartesian = hparams.get("output_format", "irreps") == "cartesian" and self.ct is not None

    def forward(self, data):
        out = self.backbone(dict(data))[OUT_FIELD_NAME]
        return self.ct.to_cartesian(out) if self.cartesian else out
