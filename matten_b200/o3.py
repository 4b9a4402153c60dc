"""
Host-side irreducible-representation algebra used to *plan* the CUDA kernels.

The reference (wengroup/matten) gets all of this from the un-vendored dependency
``e3nn==0.5.1`` (reference pyproject.toml:29).  e3nn is not a dependency of this
package: what the matten hot path needs from it -- ``Irreps`` bookkeeping
(reference src/matten/nn/utils.py:205-237), real Wigner-3j symbols, the
``normalize2mom`` constants and the ``CartesianTensor`` change of basis
(reference src/matten/utils.py:110-124) -- is re-derived here with the same
conventions so that e3nn ``state_dict``s load unchanged.  Nothing in this file
runs per batch; it produces immutable tables that are uploaded once to the GPU.
"""
from __future__ import annotations

import collections
import functools
import itertools
import math
from fractions import Fraction
from typing import List, Tuple, Union

import torch


# --------------------------------------------------------------------------- #
# Irrep / Irreps
# --------------------------------------------------------------------------- #
class Irrep(tuple):
    """(l, p) with p = +1 (even, 'e') or -1 (odd, 'o').  Ordered as a tuple, so
    ``0o < 0e < 1o < 1e`` -- the order matten's configs are written in
    (reference scripts/configs/materials_tensor.yaml:50)."""

    def __new__(cls, l: Union[int, str, "Irrep", tuple], p: int = None):
        if p is None:
            if isinstance(l, Irrep):
                return l
            if isinstance(l, str):
                s = l.strip()
                try:
                    ll = int(s[:-1])
                    pp = {"e": 1, "o": -1, "y": (-1) ** int(s[:-1])}[s[-1]]
                except Exception:
                    raise ValueError(f'unable to convert string "{l}" into an Irrep')
                l, p = ll, pp
            elif isinstance(l, tuple):
                l, p = l
        if not isinstance(l, int) or l < 0:
            raise ValueError(f"l must be a non-negative integer, got {l}")
        if p not in (-1, 1):
            raise ValueError(f"parity must be +-1, got {p}")
        return super().__new__(cls, (l, p))

    @property
    def l(self) -> int:  # noqa: E743
        return self[0]

    @property
    def p(self) -> int:
        return self[1]

    @property
    def dim(self) -> int:
        return 2 * self.l + 1

    def is_scalar(self) -> bool:
        return self.l == 0 and self.p == 1

    def __repr__(self):
        return f"{self.l}{'e' if self.p == 1 else 'o'}"

    def __mul__(self, other):
        other = Irrep(other)
        p = self.p * other.p
        for l in range(abs(self.l - other.l), self.l + other.l + 1):
            yield Irrep(l, p)

    def __rmul__(self, mul: int):
        assert isinstance(mul, int)
        return Irreps([(mul, self)])

    def __add__(self, other):
        return Irreps(self) + Irreps(other)


class _MulIr(tuple):
    def __new__(cls, mul, ir=None):
        if ir is None:
            mul, ir = mul
        assert isinstance(mul, int) and mul >= 0
        return super().__new__(cls, (mul, Irrep(ir)))

    @property
    def mul(self) -> int:
        return self[0]

    @property
    def ir(self) -> Irrep:
        return self[1]

    @property
    def dim(self) -> int:
        return self.mul * self.ir.dim

    def __repr__(self):
        return f"{self.mul}x{self.ir}"


class Irreps(tuple):
    """Direct sum of irreps, e.g. ``Irreps("32x0o+32x0e+16x1o")``."""

    def __new__(cls, irreps=None):
        if isinstance(irreps, Irreps):
            return super().__new__(cls, irreps)
        out = []
        if isinstance(irreps, Irrep):
            out.append(_MulIr(1, irreps))
        elif isinstance(irreps, str):
            try:
                if irreps.strip() != "":
                    for part in irreps.split("+"):
                        if "x" in part:
                            mul, ir = part.split("x")
                            mul = int(mul)
                        else:
                            mul, ir = 1, part
                        out.append(_MulIr(mul, Irrep(ir)))
            except Exception:
                raise ValueError(f'Unable to convert string "{irreps}" into an Irreps')
        elif irreps is None:
            pass
        else:
            for item in irreps:
                if isinstance(item, str):
                    out.append(_MulIr(1, Irrep(item)))
                elif isinstance(item, Irrep):
                    out.append(_MulIr(1, item))
                elif isinstance(item, _MulIr):
                    out.append(item)
                elif len(item) == 2:
                    mul, ir = item
                    out.append(_MulIr(int(mul), Irrep(ir)))
                else:
                    raise ValueError(f'Unable to interpret "{item}" as an irrep.')
        return super().__new__(cls, out)

    @staticmethod
    def spherical_harmonics(lmax: int, p: int = -1) -> "Irreps":
        return Irreps([(1, (l, p**l)) for l in range(lmax + 1)])

    def slices(self) -> List[slice]:
        s, i = [], 0
        for mul_ir in self:
            s.append(slice(i, i + mul_ir.dim))
            i += mul_ir.dim
        return s

    def __getitem__(self, i):
        x = super().__getitem__(i)
        if isinstance(i, slice):
            return Irreps(x)
        return x

    def __contains__(self, ir) -> bool:
        ir = Irrep(ir)
        return ir in (irrep for _, irrep in self)

    def count(self, ir) -> int:
        ir = Irrep(ir)
        return sum(mul for mul, irrep in self if ir == irrep)

    def __add__(self, irreps):
        irreps = Irreps(irreps)
        return Irreps(super().__add__(irreps))

    def __mul__(self, other):
        if isinstance(other, Irreps):
            raise NotImplementedError("Use o3.TensorProduct for this")
        return Irreps(super().__mul__(other))

    __rmul__ = __mul__

    def simplify(self) -> "Irreps":
        out = []
        for mul, ir in self:
            if out and out[-1][1] == ir:
                out[-1] = (out[-1][0] + mul, ir)
            elif mul > 0:
                out.append((mul, ir))
        return Irreps(out)

    def remove_zero_multiplicities(self) -> "Irreps":
        return Irreps([(mul, ir) for mul, ir in self if mul > 0])

    def sort(self):
        """Returns (irreps, p, inv) like e3nn: ``p[i_old] = i_new`` (the reference
        relies on this at src/matten/nn/utils.py:222-228)."""
        Ret = collections.namedtuple("sort", ["irreps", "p", "inv"])
        out = sorted([(ir, i, mul) for i, (mul, ir) in enumerate(self)])
        inv = tuple(i for _, i, _ in out)
        p = [0] * len(inv)
        for new, old in enumerate(inv):
            p[old] = new
        irreps = Irreps([(mul, ir) for ir, _, mul in out])
        return Ret(irreps, tuple(p), inv)

    @property
    def dim(self) -> int:
        return sum(mul * ir.dim for mul, ir in self)

    @property
    def num_irreps(self) -> int:
        return sum(mul for mul, _ in self)

    @property
    def ls(self) -> List[int]:
        return [ir.l for mul, ir in self for _ in range(mul)]

    @property
    def lmax(self) -> int:
        if len(self) == 0:
            raise ValueError("Cannot get lmax of empty Irreps")
        return max(self.ls)

    def __repr__(self):
        return "+".join(f"{mul_ir}" for mul_ir in self)


# --------------------------------------------------------------------------- #
# Wigner 3j in the real basis (e3nn 0.5.x conventions)
# --------------------------------------------------------------------------- #
def _f(n) -> int:
    assert n == round(n) and n >= 0
    return math.factorial(round(n))


def _su2_cg_coeff(j1, m1, j2, m2, j3, m3):
    """<j1 m1 j2 m2 | j3 m3>, Condon-Shortley, exact rationals under the root."""
    if m3 != m1 + m2:
        return 0.0
    vmin = int(max(-j1 + j2 + m3, -j1 + m1, 0))
    vmax = int(min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3))
    C = (
        (2.0 * j3 + 1.0)
        * Fraction(
            _f(j3 + j1 - j2) * _f(j3 - j1 + j2) * _f(j1 + j2 - j3) * _f(j3 + m3) * _f(j3 - m3),
            _f(j1 + j2 + j3 + 1) * _f(j1 - m1) * _f(j1 + m1) * _f(j2 - m2) * _f(j2 + m2),
        )
    ) ** 0.5
    S = 0
    for v in range(vmin, vmax + 1):
        S += (-1) ** int(v + j2 + m2) * Fraction(
            _f(j2 + j3 + m1 - v) * _f(j1 - m1 + v),
            _f(v) * _f(j3 - j1 + j2 - v) * _f(j3 + m3 - v) * _f(v + j1 - j2 - m3),
        )
    return float(C * S)


def _real_to_complex(l: int) -> torch.Tensor:
    q = torch.zeros((2 * l + 1, 2 * l + 1), dtype=torch.complex128)
    s = 1 / 2**0.5
    for m in range(-l, 0):
        q[l + m, l + abs(m)] = s
        q[l + m, l - abs(m)] = -1j * s
    q[l, l] = 1
    for m in range(1, l + 1):
        q[l + m, l + abs(m)] = (-1) ** m * s
        q[l + m, l - abs(m)] = 1j * (-1) ** m * s
    return (-1j) ** l * q


@functools.lru_cache(maxsize=None)
def _wigner_3j_f64(l1: int, l2: int, l3: int) -> torch.Tensor:
    cg = torch.zeros((2 * l1 + 1, 2 * l2 + 1, 2 * l3 + 1), dtype=torch.float64)
    for m1 in range(-l1, l1 + 1):
        for m2 in range(-l2, l2 + 1):
            if abs(m1 + m2) <= l3:
                cg[l1 + m1, l2 + m2, l3 + m1 + m2] = _su2_cg_coeff(l1, m1, l2, m2, l3, m1 + m2)
    Q1, Q2, Q3 = _real_to_complex(l1), _real_to_complex(l2), _real_to_complex(l3)
    C = torch.einsum("ij,kl,mn,ikn->jlm", Q1, Q2, torch.conj(Q3.T), cg.to(torch.complex128))
    assert torch.all(torch.abs(C.imag) < 1e-9)
    C = C.real
    return C / torch.linalg.norm(C)


def wigner_3j(l1: int, l2: int, l3: int, dtype=torch.float64) -> torch.Tensor:
    """Real-basis Wigner 3j tensor [2l1+1, 2l2+1, 2l3+1], Frobenius norm 1."""
    assert abs(l2 - l3) <= l1 <= l2 + l3
    return _wigner_3j_f64(l1, l2, l3).to(dtype).clone()


# --------------------------------------------------------------------------- #
# second-moment normalisation of activations (e3nn.math.normalize2mom)
# --------------------------------------------------------------------------- #
@functools.lru_cache(maxsize=None)
def _gauss_samples() -> torch.Tensor:
    gen = torch.Generator(device="cpu").manual_seed(0)
    return torch.randn(1_000_000, generator=gen, dtype=torch.float64)


def normalize2mom_const(f) -> float:
    """``c`` such that ``c*f(z)`` has unit second moment for z~N(0,1), estimated
    exactly the way e3nn does (1e6 fp64 samples of the CPU generator seeded 0) so
    that checkpoints trained with e3nn see the same constant.  Returns 1.0 when
    e3nn would treat the function as already normalised."""
    with torch.no_grad():
        cst = f(_gauss_samples()).pow(2).mean().pow(-0.5).item()
    if abs(cst - 1) < 1e-4:
        return 1.0
    return cst


def _ssp(x):
    return torch.nn.functional.softplus(x) - math.log(2.0)


#: activation table: name -> (callable, integer id understood by the CUDA side)
ACT_FUNCS = {
    "silu": (torch.nn.functional.silu, 1),
    "tanh": (torch.tanh, 2),
    "sigmoid": (torch.sigmoid, 3),
    "ssp": (_ssp, 4),
    "abs": (torch.abs, 5),
}


def act_parity(f) -> int:
    """+1 even, -1 odd, 0 neither (how e3nn.nn.Activation decides the output
    parity of an odd scalar)."""
    x = torch.linspace(0, 10, 256, dtype=torch.float64)
    a1, a2 = f(x), f(-x)
    if (a1 - a2).abs().max() < 1e-5:
        return 1
    if (a1 + a2).abs().max() < 1e-5:
        return -1
    return 0


# --------------------------------------------------------------------------- #
# CartesianTensor (e3nn.io.CartesianTensor + o3.ReducedTensorProducts)
# --------------------------------------------------------------------------- #
def _perm_inverse(p):
    inv = [0] * len(p)
    for i, j in enumerate(p):
        inv[j] = i
    return tuple(inv)


def _perm_compose(p1, p2):
    # (p1 . p2)(i) = p1[p2[i]]
    return tuple(p1[p2[i]] for i in range(len(p1)))


def _germinate(formula: str):
    formulas = [(-1 if f.startswith("-") else 1, f.replace("-", "")) for f in formula.split("=")]
    s0, f0 = formulas[0]
    assert s0 == 1
    for _s, f in formulas:
        if len(set(f)) != len(f) or set(f) != set(f0):
            raise RuntimeError(f"{f} is not a permutation of {f0}")
    group = {(s, tuple(f.index(i) for i in f0)) for s, f in formulas}
    while True:
        n = len(group)
        group = group.union([(s, _perm_inverse(p)) for s, p in group])
        group = group.union(
            [(s1 * s2, _perm_compose(p1, p2)) for s1, p1 in group for s2, p2 in group]
        )
        if len(group) == n:
            break
    return f0, group


def _symmetric_basis(f0: str, group, dim: int) -> torch.Tensor:
    """Orthonormal basis [d_sym, dim**rank] of the tensors invariant under the
    signed index-permutation group."""
    rank = len(f0)
    full = list(itertools.product(range(dim), repeat=rank))
    base = set()
    for x in full:
        xs = {(s, tuple(x[i] for i in p)) for s, p in group}
        if (-1, x) not in xs:
            base.add(frozenset({frozenset(xs), frozenset({(-s, y) for s, y in xs})}))
    base = sorted([sorted([sorted(xs) for xs in x]) for x in base])
    Q = torch.zeros(len(base), len(full), dtype=torch.float64)
    for i, x in enumerate(base):
        x = max(x, key=lambda xs: sum(s for s, _ in xs))
        for s, e in x:
            j = 0
            for k in e:
                j = j * dim + k
            Q[i, j] = s / len(x) ** 0.5
    return Q


def _wigner_nj(n: int, l: int = 1, p: int = -1):
    """Left-to-right coupling of n copies of the irrep (l,p) (all mul 1).
    Returns a list of (ir_out, C[ir_out.dim, d, ..., d]) sorted (stably) by irrep at
    every level of the recursion, component-normalised."""
    ir = Irrep(l, p)
    d = ir.dim
    if n == 1:
        return [(ir, torch.eye(d, dtype=torch.float64))]
    ret = []
    for ir_left, C_left in _wigner_nj(n - 1, l, p):
        for ir_out in ir_left * ir:
            C = wigner_3j(ir_out.l, ir_left.l, ir.l) * ir_out.dim**0.5
            C = torch.einsum("jk,ijl->ikl", C_left.flatten(1), C)
            C = C.reshape(ir_out.dim, *([d] * (n - 1)), d)
            ret.append((ir_out, C))
    return sorted(ret, key=lambda x: x[0])


def _orthonormalize(original: torch.Tensor, eps: float = 1e-9) -> torch.Tensor:
    final = []
    for x in original:
        for y in final:
            x = x - torch.dot(x, y) * y
        if x.norm() > 2 * eps:
            x = x / x.norm()
            x = torch.where(x.abs() < eps, torch.zeros_like(x), x)
            x = x * x[x.nonzero()[0, 0]].sign()
            final.append(x)
    if not final:
        return original.new_zeros((0, original.shape[1]))
    return torch.stack(final)


@functools.lru_cache(maxsize=None)
def _reduced_basis(formula: str, eps: float = 1e-9):
    f0, group = _germinate(formula)
    rank = len(f0)
    P = _symmetric_basis(f0, group, 3)  # [a, 3**rank]
    PP = P @ P.T
    Ps = collections.OrderedDict()
    for ir, C in _wigner_nj(rank):
        Ps.setdefault(ir, []).append(C)
    blocks, irreps_out = [], []
    for ir, bases in Ps.items():
        mul = len(bases)
        base_o3 = torch.stack(bases)  # [mul, ir.dim, 3, ..., 3]
        R = base_o3.flatten(2)  # [mul, ir.dim, 3**rank]
        R0 = R[:, 0]
        RR = R0 @ R0.T
        RP = R0 @ P.T
        prob = torch.cat([torch.cat([RR, -RP], 1), torch.cat([-RP.T, PP], 1)], 0)
        evals, evecs = torch.linalg.eigh(prob)
        X = evecs[:, evals < eps][:mul].T
        proj = X.T @ X
        for x in _orthonormalize(proj, eps):
            C = torch.einsum("u,ui...->i...", x, base_o3)
            C = C * (ir.dim / C.pow(2).sum()) ** 0.5
            blocks.append(C)
            irreps_out.append((1, ir))
    if not blocks:
        raise RuntimeError(f"formula {formula} has no symmetric tensors")
    Q = torch.cat(blocks)  # [irreps.dim, 3, ..., 3]
    return Irreps(irreps_out).simplify(), Q


class ReducedTensorProducts:
    """Holds ``change_of_basis`` [irreps_out.dim, 3, ..., 3] and ``irreps_out``."""

    def __init__(self, formula: str, dtype=None):
        self.formula = formula
        irreps_out, Q = _reduced_basis(formula)
        self.irreps_out = irreps_out
        self.change_of_basis = Q.to(dtype or torch.get_default_dtype())

    def to(self, *args, **kwargs):
        new = ReducedTensorProducts.__new__(ReducedTensorProducts)
        new.formula, new.irreps_out = self.formula, self.irreps_out
        new.change_of_basis = self.change_of_basis.to(*args, **kwargs)
        return new


class CartesianTensor(Irreps):
    """Irreps of a Cartesian tensor of polar vectors with index symmetries, e.g.
    ``"ijkl=jikl=klij"`` -> ``2x0e+2x2e+1x4e`` (reference
    src/matten/model_factory/tfn_scalar_tensor.py:47)."""

    def __new__(cls, formula: str):
        irreps_out, _ = _reduced_basis(formula)
        ret = super().__new__(cls, irreps_out)
        ret.formula = formula
        ret.indices = formula.split("=")[0].replace("-", "")
        return ret

    def reduced_tensor_products(self, dtype=None) -> ReducedTensorProducts:
        return ReducedTensorProducts(self.formula, dtype=dtype)

    def change_of_basis(self, dtype=None) -> torch.Tensor:
        """Q flattened to [irreps.dim, 3**rank]."""
        _, Q = _reduced_basis(self.formula)
        return Q.flatten(1).to(dtype or torch.get_default_dtype())


def parse_irreps_list(irreps: Irreps) -> List[Tuple[int, int, int]]:
    return [(mul, ir.l, ir.p) for mul, ir in Irreps(irreps)]
