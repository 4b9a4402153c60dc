"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Not imported by the product package.

CPU restatement (torch, fp32/fp64, autograd-capable) of the parts of the
third-party dependency ``e3nn==0.5.1`` (pinned in the reference at
pyproject.toml:29 and pretrained/20230627/conda-environment.yaml:210) that the
matten hot path calls.  e3nn is NOT vendored in /root/reference and is NOT
installable here (no network), so each function restates e3nn's published
algorithm and cites the reference call site that fixes its arguments.

PARITY UNPINNED against e3nn itself: there is no e3nn install, no golden
tensor and no checkpoint available offline.  What pins this file instead
(tests/test_oracle_*.py):
  * sympy.physics.wigner / sympy.physics.quantum.cg (third-party) for the SU(2)
    Clebsch-Gordan coefficients behind ``wigner_3j``;
  * the explicit e3nn spherical-harmonic polynomials for l<=3;
  * equivariance under Wigner-D rotations, index symmetries of the
    CartesianTensor basis, component normalisation (the property tests of the
    reference, tests/model/test_tfn_tensor.py:98-139);
  * the reference's integer KAT tests/nn/test_embedding.py:6-13.

Deliberately written in the *materialising* style of e3nn (one einsum per
instruction over dense one-hot attributes) so that it is also a fair CPU
baseline for the reference path.
"""
from __future__ import annotations

import collections
import functools
import itertools
import math
from fractions import Fraction
from math import factorial
from typing import List, Optional, Tuple

import numpy as np
import torch

# an "irreps" in this file is a plain list of (mul, l, p)
IrrepsT = List[Tuple[int, int, int]]


# ------------------------------------------------------------------ irreps --
def parse_irreps(s) -> IrrepsT:
    """e3nn.o3.Irreps(str) -> [(mul, l, p)]."""
    if not isinstance(s, str):
        return [(int(m), int(l), int(p)) for m, l, p in s]
    out = []
    for part in s.replace(" ", "").split("+"):
        if part == "":
            continue
        mul, ir = part.split("x") if "x" in part else (1, part)
        out.append((int(mul), int(ir[:-1]), 1 if ir[-1] == "e" else -1))
    return out


def irreps_str(irreps: IrrepsT) -> str:
    return "+".join(f"{m}x{l}{'e' if p == 1 else 'o'}" for m, l, p in irreps)


def irreps_dim(irreps: IrrepsT) -> int:
    return sum(m * (2 * l + 1) for m, l, _ in irreps)


def irreps_simplify(irreps: IrrepsT) -> IrrepsT:
    out = []
    for m, l, p in irreps:
        if out and out[-1][1:] == (l, p):
            out[-1] = (out[-1][0] + m, l, p)
        elif m > 0:
            out.append((m, l, p))
    return out


def irreps_sort(irreps: IrrepsT):
    """e3nn Irreps.sort(): sorted by (l, p) with p=-1 first, stable; returns
    (sorted, p) with p[i_old] = i_new."""
    order = sorted(range(len(irreps)), key=lambda i: ((irreps[i][1], irreps[i][2]), i))
    p = [0] * len(irreps)
    for new, old in enumerate(order):
        p[old] = new
    return [irreps[i] for i in order], p


def irrep_product(l1, p1, l2, p2):
    return [(l, p1 * p2) for l in range(abs(l1 - l2), l1 + l2 + 1)]


def irreps_slices(irreps: IrrepsT):
    out, i = [], 0
    for m, l, _ in irreps:
        out.append(slice(i, i + m * (2 * l + 1)))
        i += m * (2 * l + 1)
    return out


# ---------------------------------------------------------------- wigner 3j --
def su2_clebsch_gordan_coeff(j1, m1, j2, m2, j3, m3) -> float:
    """e3nn.o3._wigner._su2_clebsch_gordan_coeff."""
    if m3 != m1 + m2:
        return 0.0
    vmin = int(max([-j1 + j2 + m3, -j1 + m1, 0]))
    vmax = int(min([j2 + j3 + m1, j3 - j1 + j2, j3 + m3]))

    def f(n):
        return factorial(round(n))

    C = (
        (2.0 * j3 + 1.0)
        * Fraction(
            f(j3 + j1 - j2) * f(j3 - j1 + j2) * f(j1 + j2 - j3) * f(j3 + m3) * f(j3 - m3),
            f(j1 + j2 + j3 + 1) * f(j1 - m1) * f(j1 + m1) * f(j2 - m2) * f(j2 + m2),
        )
    ) ** 0.5
    S = 0
    for v in range(vmin, vmax + 1):
        S += (-1) ** int(v + j2 + m2) * Fraction(
            f(j2 + j3 + m1 - v) * f(j1 - m1 + v),
            f(v) * f(j3 - j1 + j2 - v) * f(j3 + m3 - v) * f(v + j1 - j2 - m3),
        )
    return float(C * S)


def change_basis_real_to_complex(l: int) -> np.ndarray:
    """e3nn.o3._wigner.change_basis_real_to_complex."""
    q = np.zeros((2 * l + 1, 2 * l + 1), dtype=np.complex128)
    for m in range(-l, 0):
        q[l + m, l + abs(m)] = 1 / 2**0.5
        q[l + m, l - abs(m)] = -1j / 2**0.5
    q[l, l] = 1
    for m in range(1, l + 1):
        q[l + m, l + abs(m)] = (-1) ** m / 2**0.5
        q[l + m, l - abs(m)] = 1j * (-1) ** m / 2**0.5
    return (-1j) ** l * q


@functools.lru_cache(maxsize=None)
def _wigner_3j_np(l1: int, l2: int, l3: int) -> np.ndarray:
    """e3nn.o3._wigner._so3_clebsch_gordan."""
    mat = np.zeros((2 * l1 + 1, 2 * l2 + 1, 2 * l3 + 1))
    for m1 in range(-l1, l1 + 1):
        for m2 in range(-l2, l2 + 1):
            if abs(m1 + m2) <= l3:
                mat[l1 + m1, l2 + m2, l3 + m1 + m2] = su2_clebsch_gordan_coeff(
                    l1, m1, l2, m2, l3, m1 + m2
                )
    Q1 = change_basis_real_to_complex(l1)
    Q2 = change_basis_real_to_complex(l2)
    Q3 = change_basis_real_to_complex(l3)
    C = np.einsum("ij,kl,mn,ikn->jlm", Q1, Q2, np.conj(Q3.T), mat.astype(np.complex128))
    assert np.all(np.abs(C.imag) < 1e-5)
    C = C.real
    return C / np.linalg.norm(C)


def wigner_3j(l1, l2, l3, dtype=torch.float64) -> torch.Tensor:
    assert abs(l2 - l3) <= l1 <= l2 + l3
    return torch.from_numpy(_wigner_3j_np(l1, l2, l3).copy()).to(dtype)


# ------------------------------------------------ rotations (for the tests) --
def so3_generators(l: int) -> np.ndarray:
    """e3nn.o3._wigner.so3_generators: real generators in e3nn's basis."""
    m = np.arange(-l, l)
    raising = np.diag(-np.sqrt(l * (l + 1) - m * (m + 1)), k=-1)
    m = np.arange(-l + 1, l + 1)
    lowering = np.diag(np.sqrt(l * (l + 1) - m * (m - 1)), k=1)
    m = np.arange(-l, l + 1)
    X = np.stack(
        [
            0.5 * (raising + lowering),  # x (usually)
            np.diag(1j * m),  # z (usually)
            -0.5j * (raising - lowering),  # -y (usually)
        ],
        axis=0,
    )
    Q = change_basis_real_to_complex(l)
    X = np.conj(Q.T) @ X @ Q
    assert np.all(np.abs(X.imag) < 1e-5)
    return X.real


def _expm(A: np.ndarray) -> np.ndarray:
    import scipy.linalg

    return scipy.linalg.expm(A)


def wigner_D(l: int, alpha: float, beta: float, gamma: float) -> np.ndarray:
    """e3nn.o3.wigner_D: D^l(R) for R = Ry(alpha) Rx(beta) Ry(gamma)."""
    X = so3_generators(l)
    return _expm(alpha * X[1]) @ _expm(beta * X[0]) @ _expm(gamma * X[1])


def angles_to_matrix(alpha, beta, gamma) -> np.ndarray:
    """e3nn.o3.angles_to_matrix: Ry(alpha) Rx(beta) Ry(gamma)."""

    def Ry(a):
        c, s = math.cos(a), math.sin(a)
        return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])

    def Rx(a):
        c, s = math.cos(a), math.sin(a)
        return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])

    return Ry(alpha) @ Rx(beta) @ Ry(gamma)


# ------------------------------------------------------ spherical harmonics --
def spherical_harmonics(lmax: int, vec: torch.Tensor, normalize: bool = True,
                        normalization: str = "component") -> torch.Tensor:
    """e3nn.o3.SphericalHarmonics(Irreps.spherical_harmonics(lmax), normalize,
    normalization) as constructed at reference src/matten/nn/_nequip.py:167-169.

    e3nn ships machine-generated polynomials; they are the normalised recursion
    Y_{l+1} ~ w3j(l+1,1,l) . (Y_1 (x) Y_l), Y_1 = (x, y, z), with Y_l(e_y) = e_{m=0}
    ("norm" normalisation), then scaled per normalisation.  l<=3 is written out
    exactly as e3nn's generated code; l>=4 uses the recursion (checked against the
    explicit forms for l<=3 in tests/test_oracle_e3nn.py)."""
    if normalize:
        vec = torch.nn.functional.normalize(vec, dim=-1)
    x, y, z = vec[..., 0], vec[..., 1], vec[..., 2]
    sh = [torch.ones_like(x)[..., None]]
    if lmax >= 1:
        sh.append(torch.stack([x, y, z], -1))
    if lmax >= 2:
        y2 = y.pow(2)
        x2z2 = x.pow(2) + z.pow(2)
        sh_2_0 = math.sqrt(3.0) * x * z
        sh_2_1 = math.sqrt(3.0) * x * y
        sh_2_2 = y2 - 0.5 * x2z2
        sh_2_3 = math.sqrt(3.0) * y * z
        sh_2_4 = math.sqrt(3.0) / 2.0 * (z.pow(2) - x.pow(2))
        sh.append(torch.stack([sh_2_0, sh_2_1, sh_2_2, sh_2_3, sh_2_4], -1))
    if lmax >= 3:
        sh_3_0 = (1 / 6) * math.sqrt(30) * (sh_2_0 * z + sh_2_4 * x)
        sh_3_1 = math.sqrt(5) * sh_2_0 * y
        sh_3_2 = (1 / 4) * math.sqrt(6) * (4.0 * y2 - x2z2) * x
        sh_3_3 = (1 / 2) * y * (2.0 * y2 - 3.0 * x2z2)
        sh_3_4 = (1 / 4) * math.sqrt(6) * z * (4.0 * y2 - x2z2)
        sh_3_5 = math.sqrt(5) * sh_2_4 * y
        sh_3_6 = (1 / 6) * math.sqrt(30) * (sh_2_4 * z - sh_2_0 * x)
        sh.append(torch.stack([sh_3_0, sh_3_1, sh_3_2, sh_3_3, sh_3_4, sh_3_5, sh_3_6], -1))
    for l in range(3, lmax):
        sh.append(_sh_recursion_step(l, sh[1], sh[l]))
    if normalization == "component":
        sh = [s * math.sqrt(2 * l + 1) for l, s in enumerate(sh)]
    elif normalization == "integral":
        sh = [s * math.sqrt(2 * l + 1) / math.sqrt(4 * math.pi) for l, s in enumerate(sh)]
    elif normalization != "norm":
        raise ValueError(normalization)
    return torch.cat(sh[: lmax + 1], -1)


@functools.lru_cache(maxsize=None)
def _sh_recursion_coeff(l: int):
    """w3j(l+1, 1, l) scaled so that the recursion maps unit-norm Y_l to unit-norm
    Y_{l+1} with Y_{l+1}(e_y)[m=0] = +1."""
    C = _wigner_3j_np(l + 1, 1, l)
    y1 = np.array([0.0, 1.0, 0.0])
    yl = np.zeros(2 * l + 1)
    yl[l] = 1.0
    v = np.einsum("kij,i,j->k", C, y1, yl)
    return C / v[l + 1]


def _sh_recursion_step(l: int, y1: torch.Tensor, yl: torch.Tensor) -> torch.Tensor:
    C = torch.from_numpy(_sh_recursion_coeff(l)).to(yl.dtype)
    return torch.einsum("kij,...i,...j->...k", C, y1, yl)


def spherical_harmonics_recursive(lmax: int, vec: torch.Tensor) -> torch.Tensor:
    """Pure recursion from l=1 ('norm' normalisation, normalised input)."""
    vec = torch.nn.functional.normalize(vec, dim=-1)
    sh = [torch.ones_like(vec[..., :1]), vec]
    for l in range(1, lmax):
        sh.append(_sh_recursion_step(l, sh[1], sh[l]))
    return torch.cat(sh[: lmax + 1], -1)


# ------------------------------------------------------------ radial basis --
def soft_one_hot_linspace_bessel(x: torch.Tensor, start: float, end: float, number: int,
                                 cutoff: bool = True) -> torch.Tensor:
    """e3nn.math.soft_one_hot_linspace(basis='bessel') as called at reference
    src/matten/nn/embedding.py:189-196."""
    x = x[..., None] - start
    c = end - start
    bessel_roots = torch.arange(1, number + 1, dtype=x.dtype, device=x.device) * math.pi
    out = math.sqrt(2 / c) * torch.sin(bessel_roots * x / c) / x
    if not cutoff:
        return out
    return out * ((x / c) < 1) * (0 < x)


# ------------------------------------------------------------ activations --
@functools.lru_cache(maxsize=None)
def _normal_samples():
    gen = torch.Generator(device="cpu").manual_seed(0)
    return torch.randn(1_000_000, generator=gen, dtype=torch.float64)


class normalize2mom(torch.nn.Module):
    """e3nn.math.normalize2mom."""

    def __init__(self, f):
        super().__init__()
        with torch.no_grad():
            cst = f(_normal_samples()).pow(2).mean().pow(-0.5).item()
        self._is_id = abs(cst - 1) < 1e-4
        self.f = f
        self.cst = cst

    def forward(self, x):
        if self._is_id:
            return self.f(x)
        return self.f(x).mul(self.cst)


class FullyConnectedNet(torch.nn.Module):
    """e3nn.nn.FullyConnectedNet(hs, act) with variance_in = variance_out = 1 and
    out_act=False, as built at reference src/matten/nn/utils.py:246-251.
    Parameters: layer{i}.weight [h_in, h_out] ~ N(0,1)."""

    def __init__(self, hs, act=None):
        super().__init__()
        self.hs = list(hs)
        self.act = normalize2mom(act) if act is not None else None
        self.n = len(self.hs) - 1
        for i, (h1, h2) in enumerate(zip(self.hs, self.hs[1:])):
            layer = torch.nn.Module()
            layer.weight = torch.nn.Parameter(torch.randn(h1, h2))
            setattr(self, f"layer{i}", layer)

    def forward(self, x):
        for i in range(self.n):
            w = getattr(self, f"layer{i}").weight
            x = x @ (w / (w.shape[0]) ** 0.5)
            if i < self.n - 1 and self.act is not None:
                x = self.act(x)
        return x


# ---------------------------------------------------------- tensor product --
Instruction = collections.namedtuple(
    "Instruction", "i_in1 i_in2 i_out connection_mode has_weight path_weight path_shape"
)


class TensorProduct(torch.nn.Module):
    """e3nn.o3.TensorProduct with irrep_normalization='component',
    path_normalization='element', in/out variances 1.  Supports the connection
    modes matten uses: 'uvu' (nn/utils.py:213), 'uvw' (FullyConnectedTensorProduct)
    and 'uuu' (inside e3nn.nn.Gate).  One einsum per instruction, like e3nn's
    codegen ("zuvij" outer product then contraction with the w3j)."""

    def __init__(self, irreps_in1, irreps_in2, irreps_out, instructions,
                 internal_weights=False, shared_weights=False):
        super().__init__()
        self.irreps_in1 = parse_irreps(irreps_in1)
        self.irreps_in2 = parse_irreps(irreps_in2)
        self.irreps_out = parse_irreps(irreps_out)
        ins = []
        for t in instructions:
            i1, i2, io, mode, hw = t[:5]
            pw = t[5] if len(t) > 5 else 1.0
            m1, m2, mo = self.irreps_in1[i1][0], self.irreps_in2[i2][0], self.irreps_out[io][0]
            shape = {"uvw": (m1, m2, mo), "uvu": (m1, m2), "uvv": (m1, m2), "uuw": (m1, mo),
                     "uuu": (m1,), "uvuv": (m1, m2)}[mode]
            ins.append(Instruction(i1, i2, io, mode, hw, pw, shape))

        def num_elements(i):
            m1, m2 = self.irreps_in1[i.i_in1][0], self.irreps_in2[i.i_in2][0]
            return {"uvw": m1 * m2, "uvu": m2, "uvv": m1, "uuw": m1, "uuu": 1, "uvuv": 1}[
                i.connection_mode]

        out = []
        for i in ins:
            alpha = 2 * self.irreps_out[i.i_out][1] + 1  # component
            x = sum(num_elements(j) for j in ins if j.i_out == i.i_out)  # element
            if x > 0:
                alpha /= x
            alpha *= i.path_weight
            out.append(i._replace(path_weight=math.sqrt(alpha)))
        self.instructions = out
        self.weight_numel = sum(int(np.prod(i.path_shape)) for i in out if i.has_weight)
        self.internal_weights = internal_weights
        self.shared_weights = shared_weights
        if internal_weights and self.weight_numel > 0:
            assert shared_weights
            self.weight = torch.nn.Parameter(torch.randn(self.weight_numel))
        else:
            self.register_buffer("weight", torch.Tensor())
        mask = []
        for io, (m, l, p) in enumerate(self.irreps_out):
            has = any(i.i_out == io and i.path_weight != 0 and 0 not in i.path_shape for i in out)
            mask.append(torch.ones(m * (2 * l + 1)) if has else torch.zeros(m * (2 * l + 1)))
        self.register_buffer("output_mask", torch.cat(mask) if mask else torch.ones(0))

    def forward(self, x1, x2, weight: Optional[torch.Tensor] = None):
        if weight is None:
            weight = self.weight
        lead = x1.shape[:-1]
        x1 = x1.reshape(-1, x1.shape[-1])
        x2 = x2.reshape(-1, x2.shape[-1])
        Z = x1.shape[0]
        if self.weight_numel > 0:
            if self.shared_weights:
                weight = weight.reshape(self.weight_numel)
            else:
                weight = weight.reshape(-1, self.weight_numel)
        s1, s2 = irreps_slices(self.irreps_in1), irreps_slices(self.irreps_in2)
        outs = [None] * len(self.irreps_out)
        off = 0
        for ins in self.instructions:
            m1, l1, _ = self.irreps_in1[ins.i_in1]
            m2, l2, _ = self.irreps_in2[ins.i_in2]
            mo, lo, _ = self.irreps_out[ins.i_out]
            a = x1[:, s1[ins.i_in1]].reshape(Z, m1, 2 * l1 + 1)
            b = x2[:, s2[ins.i_in2]].reshape(Z, m2, 2 * l2 + 1)
            w = None
            if ins.has_weight:
                n = int(np.prod(ins.path_shape))
                if self.shared_weights:
                    w = weight[off:off + n].reshape(ins.path_shape)
                else:
                    w = weight[:, off:off + n].reshape((-1,) + ins.path_shape)
                off += n
            if m1 * m2 * mo == 0:
                continue
            C = wigner_3j(l1, l2, lo, dtype=x1.dtype).to(x1.device)
            xx = torch.einsum("zui,zvj->zuvij", a, b)
            z = "" if self.shared_weights else "z"
            mode = ins.connection_mode
            if mode == "uvw":
                r = torch.einsum(f"{z}uvw,ijk,zuvij->zwk", w, C, xx)
            elif mode == "uvu":
                if w is not None:
                    r = torch.einsum(f"{z}uv,ijk,zuvij->zuk", w, C, xx)
                else:
                    r = torch.einsum("ijk,zuvij->zuk", C, xx)
            elif mode == "uuu":
                if w is not None:
                    r = torch.einsum(f"{z}u,ijk,zuuij->zuk", w, C, xx)
                else:
                    r = torch.einsum("ijk,zuuij->zuk", C, xx)
            else:
                raise NotImplementedError(mode)
            r = ins.path_weight * r.reshape(Z, mo * (2 * lo + 1))
            outs[ins.i_out] = r if outs[ins.i_out] is None else outs[ins.i_out] + r
        cols = []
        for io, (m, l, _) in enumerate(self.irreps_out):
            cols.append(outs[io] if outs[io] is not None else x1.new_zeros(Z, m * (2 * l + 1)))
        out = torch.cat(cols, -1) if cols else x1.new_zeros(Z, 0)
        return out.reshape(lead + (out.shape[-1],))


class FullyConnectedTensorProduct(TensorProduct):
    """e3nn.o3.FullyConnectedTensorProduct (reference src/matten/nn/conv.py:59,77,84)."""

    def __init__(self, irreps_in1, irreps_in2, irreps_out):
        i1, i2, io = parse_irreps(irreps_in1), parse_irreps(irreps_in2), parse_irreps(irreps_out)
        instr = [
            (a, b, c, "uvw", True, 1.0)
            for a, (_, l1, p1) in enumerate(i1)
            for b, (_, l2, p2) in enumerate(i2)
            for c, (_, l3, p3) in enumerate(io)
            if (l3, p3) in irrep_product(l1, p1, l2, p2)
        ]
        super().__init__(i1, i2, io, instr, internal_weights=True, shared_weights=True)


class ElementwiseTensorProduct(TensorProduct):
    """e3nn.o3.ElementwiseTensorProduct (used by Gate)."""

    def __init__(self, irreps_in1, irreps_in2):
        a = irreps_simplify(parse_irreps(irreps_in1))
        b = irreps_simplify(parse_irreps(irreps_in2))
        assert sum(m for m, _, _ in a) == sum(m for m, _, _ in b)
        a, b = list(a), list(b)
        i = 0
        while i < len(a):
            m1, l1, p1 = a[i]
            m2, l2, p2 = b[i]
            if m1 < m2:
                b[i] = (m1, l2, p2)
                b.insert(i + 1, (m2 - m1, l2, p2))
            if m2 < m1:
                a[i] = (m2, l1, p1)
                a.insert(i + 1, (m1 - m2, l1, p1))
            i += 1
        out, instr = [], []
        for i, ((m, l1, p1), (m_2, l2, p2)) in enumerate(zip(a, b)):
            assert m == m_2
            for l, p in irrep_product(l1, p1, l2, p2):
                instr.append((i, i, len(out), "uuu", False))
                out.append((m, l, p))
        super().__init__(a, b, out, instr)


class Linear(torch.nn.Module):
    """e3nn.o3.Linear(irreps_in, irreps_out) without biases (reference
    src/matten/nn/nodewise.py:111, model_factory/tfn_scalar_tensor.py:50)."""

    def __init__(self, irreps_in, irreps_out):
        super().__init__()
        self.irreps_in = parse_irreps(irreps_in)
        self.irreps_out = parse_irreps(irreps_out)
        self.instr = [
            (a, b)
            for a, (_, l1, p1) in enumerate(self.irreps_in)
            for b, (_, l2, p2) in enumerate(self.irreps_out)
            if (l1, p1) == (l2, p2)
        ]
        self.path_weight = []
        for a, b in self.instr:
            fan = sum(self.irreps_in[a2][0] for a2, b2 in self.instr if b2 == b)
            self.path_weight.append(1.0 / math.sqrt(fan))
        self.weight_numel = sum(self.irreps_in[a][0] * self.irreps_out[b][0] for a, b in self.instr)
        self.weight = torch.nn.Parameter(torch.randn(self.weight_numel))

    def forward(self, x):
        lead = x.shape[:-1]
        x = x.reshape(-1, x.shape[-1])
        Z = x.shape[0]
        si = irreps_slices(self.irreps_in)
        outs = [None] * len(self.irreps_out)
        off = 0
        for (a, b), pw in zip(self.instr, self.path_weight):
            mi, l, _ = self.irreps_in[a]
            mo = self.irreps_out[b][0]
            w = self.weight[off:off + mi * mo].reshape(mi, mo)
            off += mi * mo
            r = pw * torch.einsum("uw,zui->zwi", w, x[:, si[a]].reshape(Z, mi, 2 * l + 1))
            r = r.reshape(Z, -1)
            outs[b] = r if outs[b] is None else outs[b] + r
        cols = [outs[b] if outs[b] is not None else x.new_zeros(Z, m * (2 * l + 1))
                for b, (m, l, _) in enumerate(self.irreps_out)]
        out = torch.cat(cols, -1)
        return out.reshape(lead + (out.shape[-1],))


# --------------------------------------------------------------------- gate --
def _act_parity(f) -> int:
    x = torch.linspace(0, 10, 256, dtype=torch.float64)
    a1, a2 = f(x), f(-x)
    if (a1 - a2).abs().max() < 1e-5:
        return 1
    if (a1 + a2).abs().max() < 1e-5:
        return -1
    return 0


class Activation(torch.nn.Module):
    """e3nn.nn.Activation on scalar irreps, each act wrapped in normalize2mom."""

    def __init__(self, irreps_in, acts):
        super().__init__()
        self.irreps_in = parse_irreps(irreps_in)
        assert len(self.irreps_in) == len(acts)
        self._acts = [normalize2mom(a) if a is not None else None for a in acts]
        out = []
        for (m, l, p), a in zip(self.irreps_in, self._acts):
            if a is not None:
                assert l == 0, "Activation: cannot apply an activation function to a non-scalar input."
                pa = _act_parity(a)
                p_out = pa if p == -1 else p
                if p_out == 0:
                    raise ValueError("Activation: the parity is violated!")
                out.append((m, 0, p_out))
            else:
                out.append((m, l, p))
        self.irreps_out = out

    def forward(self, x):
        cols, i = [], 0
        for (m, l, _), a in zip(self.irreps_in, self._acts):
            d = m * (2 * l + 1)
            blk = x[..., i:i + d]
            cols.append(a(blk) if a is not None else blk)
            i += d
        return torch.cat(cols, -1) if cols else x


class Gate(torch.nn.Module):
    """e3nn.nn.Gate (reference src/matten/nn/utils.py:134-140)."""

    def __init__(self, irreps_scalars, act_scalars, irreps_gates, act_gates, irreps_gated):
        super().__init__()
        s, g, t = parse_irreps(irreps_scalars), parse_irreps(irreps_gates), parse_irreps(irreps_gated)
        assert all(l == 0 for _, l, _ in s) and all(l == 0 for _, l, _ in g)
        assert sum(m for m, _, _ in g) == sum(m for m, _, _ in t)
        self.s, self.g, self.t = s, g, t
        self.irreps_in = irreps_simplify(s + g + t)
        self.act_scalars = Activation(s, act_scalars)
        self.act_gates = Activation(g, act_gates)
        self.mul = ElementwiseTensorProduct(t, self.act_gates.irreps_out)
        self.irreps_out = self.act_scalars.irreps_out + self.mul.irreps_out

    def forward(self, x):
        ds, dg = irreps_dim(self.s), irreps_dim(self.g)
        scalars, gates, gated = x[..., :ds], x[..., ds:ds + dg], x[..., ds + dg:]
        scalars = self.act_scalars(scalars)
        if dg > 0:
            gates = self.act_gates(gates)
            gated = self.mul(gated, gates)
            return torch.cat([scalars, gated], -1)
        return scalars


# ----------------------------------------------------------- norm activation --
class NormActivation(torch.nn.Module):
    """e3nn.nn.NormActivation (e3nn 0.5.x ``nn/_normact.py``): o3.Norm(irreps, squared=True) is the plain sum of
    squares per channel, clamped from below at epsilon^2 and square-rooted; the scalar nonlinearity acts on the norm
    (plus an optional per-channel bias), is divided by the norm when ``normalize`` and multiplies the channel through
    ElementwiseTensorProduct("Nx0e", irreps) -- a plain scalar * vector product under component normalisation.
    The reference builds it with normalize=True, epsilon=1e-8, bias=False (src/matten/nn/utils.py:142-150)."""

    def __init__(self, irreps_in, scalar_nonlinearity, normalize=True, epsilon=None, bias=False):
        super().__init__()
        self.irreps_in = self.irreps_out = parse_irreps(irreps_in)
        if epsilon is None and normalize:
            epsilon = 1e-8
        elif epsilon is not None and not normalize:
            raise ValueError("epsilon and normalize = False don't make sense together")
        self._eps_squared = epsilon * epsilon if epsilon is not None else 0.0
        self.scalar_nonlinearity, self.normalize, self.bias = scalar_nonlinearity, normalize, bias
        n = sum(m for m, _, _ in self.irreps_in)
        if bias:
            self.biases = torch.nn.Parameter(torch.zeros(n))

    def forward(self, features):
        norms = []
        for (m, l, p), sl in zip(self.irreps_in, irreps_slices(self.irreps_in)):
            f = features[..., sl].reshape(features.shape[:-1] + (m, 2 * l + 1))
            norms.append(f.pow(2).sum(-1))
        norms = torch.cat(norms, -1)
        if self._eps_squared > 0:
            norms = torch.where(norms < self._eps_squared, torch.full_like(norms, self._eps_squared), norms).sqrt()
        else:
            norms = norms.sqrt()
        arg = norms + self.biases if self.bias else norms
        scalings = self.scalar_nonlinearity(arg)
        if self.normalize:
            scalings = scalings / norms
        out, i = [], 0
        for (m, l, p), sl in zip(self.irreps_in, irreps_slices(self.irreps_in)):
            f = features[..., sl].reshape(features.shape[:-1] + (m, 2 * l + 1))
            out.append((f * scalings[..., i:i + m, None]).reshape(features.shape[:-1] + (m * (2 * l + 1),)))
            i += m
        return torch.cat(out, -1)


# ---------------------------------------------------------------- batchnorm --
class BatchNorm(torch.nn.Module):
    """e3nn.nn.BatchNorm(irreps) defaults: eps 1e-5, momentum 0.1, affine,
    reduce='mean', instance=False, normalization='component' (reference
    src/matten/nn/utils.py:418).  Only 0e blocks are mean-centred and biased
    (``Irrep.is_scalar`` means l==0 and p==+1)."""

    def __init__(self, irreps, eps=1e-5, momentum=0.1):
        super().__init__()
        self.irreps = parse_irreps(irreps)
        self.eps, self.momentum = eps, momentum
        ns = sum(m for m, l, p in self.irreps if l == 0 and p == 1)
        nf = sum(m for m, _, _ in self.irreps)
        self.register_buffer("running_mean", torch.zeros(ns))
        self.register_buffer("running_var", torch.ones(nf))
        self.weight = torch.nn.Parameter(torch.ones(nf))
        self.bias = torch.nn.Parameter(torch.zeros(ns))

    def forward(self, x):
        lead = x.shape[:-1]
        dim = x.shape[-1]
        x = x.reshape(-1, 1, dim)  # e3nn: [batch, sample, dim]
        new_means, new_vars, fields = [], [], []
        ix = irm = irv = iw = ib = 0
        for m, l, p in self.irreps:
            d = 2 * l + 1
            field = x[:, :, ix:ix + m * d].reshape(x.shape[0], -1, m, d)
            ix += m * d
            if l == 0 and p == 1:
                if self.training:
                    mean = field.mean([0, 1]).reshape(m)
                    new_means.append((1 - self.momentum) * self.running_mean[irm:irm + m]
                                     + self.momentum * mean.detach())
                else:
                    mean = self.running_mean[irm:irm + m]
                irm += m
                field = field - mean.reshape(-1, 1, m, 1)
            if self.training:
                norm = field.pow(2).mean(3).mean(1).mean(0)
                new_vars.append((1 - self.momentum) * self.running_var[irv:irv + m]
                                + self.momentum * norm.detach())
            else:
                norm = self.running_var[irv:irv + m]
            irv += m
            norm = (norm + self.eps).pow(-0.5) * self.weight[iw:iw + m]
            iw += m
            field = field * norm.reshape(-1, 1, m, 1)
            if l == 0 and p == 1:
                field = field + self.bias[ib:ib + m].reshape(m, 1)
                ib += m
            fields.append(field.reshape(x.shape[0], -1, m * d))
        if self.training:
            with torch.no_grad():
                if new_means:
                    self.running_mean.copy_(torch.cat(new_means))
                if new_vars:
                    self.running_var.copy_(torch.cat(new_vars))
        return torch.cat(fields, 2).reshape(lead + (dim,))


# ------------------------------------------------------------------ scatter --
def scatter(src: torch.Tensor, index: torch.Tensor, dim_size: Optional[int] = None,
            reduce: str = "sum") -> torch.Tensor:
    """torch_scatter.scatter(src, index, dim=0, dim_size, reduce) (reference
    src/matten/nn/conv.py:114, nn/nodewise.py:144)."""
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    idx = index.reshape(-1, *([1] * (src.dim() - 1))).expand_as(src)
    out = src.new_zeros((dim_size,) + src.shape[1:])
    if reduce in ("sum", "add"):
        return out.scatter_add_(0, idx, src)
    if reduce == "mean":
        out.scatter_add_(0, idx, src)
        cnt = torch.zeros(dim_size, dtype=src.dtype, device=src.device).scatter_add_(
            0, index, torch.ones_like(index, dtype=src.dtype)).clamp_(min=1)
        return out / cnt.reshape(-1, *([1] * (src.dim() - 1)))
    if reduce in ("min", "max"):
        return out.scatter_reduce_(0, idx, src, "a" + reduce, include_self=False)
    raise ValueError(reduce)


# ----------------------------------------------------------- cartesian tensor --
def _perm_inverse(p):
    inv = [0] * len(p)
    for i, j in enumerate(p):
        inv[j] = i
    return tuple(inv)


def _perm_compose(p1, p2):
    return tuple(p1[p2[i]] for i in range(len(p1)))


def germinate_formulas(formula: str):
    """e3nn.util.germinate_formulas."""
    formulas = [(-1 if f.startswith("-") else 1, f.replace("-", "")) for f in formula.split("=")]
    s0, f0 = formulas[0]
    assert s0 == 1
    for _s, f in formulas:
        if len(set(f)) != len(f) or set(f) != set(f0):
            raise RuntimeError(f"{f} is not a permutation of {f0}")
    formulas = {(s, tuple(f.index(i) for i in f0)) for s, f in formulas}
    while True:
        n = len(formulas)
        formulas = formulas.union([(s, _perm_inverse(p)) for s, p in formulas])
        formulas = formulas.union(
            [(s1 * s2, _perm_compose(p1, p2)) for s1, p1 in formulas for s2, p2 in formulas])
        if len(formulas) == n:
            break
    return f0, formulas


def reduce_permutation(f0, formulas, dim: int) -> np.ndarray:
    """e3nn.util.reduce_permutation -> Q[d_sym, dim, ..., dim] (orthonormal rows)."""
    full_base = list(itertools.product(*(range(dim) for _ in f0)))
    base = set()
    for x in full_base:
        xs = {(s, tuple(x[i] for i in p)) for s, p in formulas}
        if not (-1, x) in xs:
            base.add(frozenset({frozenset(xs), frozenset({(-s, x) for s, x in xs})}))
    base = sorted([sorted([sorted(xs) for xs in x]) for x in base])
    Q = np.zeros((len(base), len(full_base)))
    for i, x in enumerate(base):
        x = max(x, key=lambda xs: sum(s for s, x in xs))
        for s, e in x:
            j = 0
            for k in e:
                j = j * dim + k
            Q[i, j] = s / len(x) ** 0.5
    return Q.reshape((len(base),) + (dim,) * len(f0))


def _wigner_nj(ls_ps, dtype=np.float64):
    """e3nn.o3._reduce._wigner_nj for a list of single irreps [(l,p), ...] (each
    'irreps' has one irrep of multiplicity 1, as CartesianTensor uses '1o')."""
    if len(ls_ps) == 1:
        (l, p), = ls_ps
        return [((l, p), np.eye(2 * l + 1))]
    *left, (lr, pr) = ls_ps
    ret = []
    for (ll, pl), C_left in _wigner_nj(left):
        for lo, po in irrep_product(ll, pl, lr, pr):
            C = _wigner_3j_np(lo, ll, lr) * (2 * lo + 1) ** 0.5
            C = np.einsum("jk,ijl->ikl", C_left.reshape(C_left.shape[0], -1), C)
            C = C.reshape((2 * lo + 1,) + tuple(2 * l + 1 for l, _ in left) + (2 * lr + 1,))
            ret.append(((lo, po), C))
    return sorted(ret, key=lambda x: x[0])


def orthonormalize(original: np.ndarray, eps: float = 1e-9) -> np.ndarray:
    """e3nn.math.orthonormalize (only the orthonormal rows are returned)."""
    final = []
    for x in original:
        x = x.copy()
        for y in final:
            x = x - np.dot(x, y) * y
        if np.linalg.norm(x) > 2 * eps:
            x = x / np.linalg.norm(x)
            x[np.abs(x) < eps] = 0
            x = x * np.sign(x[np.nonzero(x)[0][0]])
            final.append(x)
    return np.stack(final) if final else np.zeros((0, original.shape[1]))


@functools.lru_cache(maxsize=None)
def reduced_tensor_products(formula: str, eps: float = 1e-9):
    """e3nn.o3.ReducedTensorProducts(formula, **{i: '1o'}) ->
    (irreps_out, change_of_basis[irreps_out.dim, 3, ..., 3]) (fp64)."""
    f0, formulas = germinate_formulas(formula)
    P = reduce_permutation(f0, formulas, 3)
    P = P.reshape(P.shape[0], -1)
    PP = P @ P.T
    Ps = collections.OrderedDict()
    for ir, base in _wigner_nj([(1, -1)] * len(f0)):
        Ps.setdefault(ir, []).append(base)
    change_of_basis, irreps_out = [], []
    for ir in Ps:
        mul = len(Ps[ir])
        base_o3 = np.stack(Ps[ir])
        R = base_o3.reshape(mul, 2 * ir[0] + 1, -1)
        RR = R[:, 0] @ R[:, 0].T
        RP = R[:, 0] @ P.T
        prob = np.block([[RR, -RP], [-RP.T, PP]])
        eigenvalues, eigenvectors = np.linalg.eigh(prob)
        X = eigenvectors[:, eigenvalues < eps][:mul].T
        proj = X.T @ X
        for x in orthonormalize(proj, eps):
            C = np.einsum("u,ui...->i...", x, base_o3)
            C = ((2 * ir[0] + 1) / (C**2).sum()) ** 0.5 * C
            change_of_basis.append(C)
            irreps_out.append((1, ir[0], ir[1]))
    return irreps_simplify(irreps_out), np.concatenate(change_of_basis)


class CartesianTensor:
    """e3nn.io.CartesianTensor(formula) with indices of irrep 1o."""

    def __init__(self, formula: str):
        self.formula = formula
        self.indices = formula.split("=")[0].replace("-", "")
        self.irreps, Q = reduced_tensor_products(formula)
        self.Q = torch.from_numpy(Q.copy())

    def from_cartesian(self, data: torch.Tensor) -> torch.Tensor:
        Q = self.Q.to(data.dtype).flatten(1)
        return data.flatten(-len(self.indices)) @ Q.T

    def to_cartesian(self, data: torch.Tensor) -> torch.Tensor:
        Q = self.Q.to(data.dtype)
        out = data @ Q.flatten(1)
        return out.reshape(tuple(data.shape[:-1]) + tuple(Q.shape[1:]))
