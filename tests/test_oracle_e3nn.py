"""Pins of the oracle's e3nn restatement (oracle/e3nn_restated.py).  e3nn itself is not
installable offline, so these are: a third-party check (sympy) of the SU(2) Clebsch-Gordan
coefficients, e3nn's explicit spherical-harmonic polynomials for l <= 3 against the recursion,
Wigner-D equivariance of everything, and the analytic identities listed in SURVEY.md section 8c."""
import math

import numpy as np
import pytest
import torch

from oracle import e3nn_restated as E

ANGLES = (0.3, 1.1, -0.7)


def test_su2_cg_against_sympy():
    from sympy import S
    from sympy.physics.quantum.cg import CG

    worst = 0.0
    for j1 in range(4):
        for j2 in range(4):
            for j3 in range(abs(j1 - j2), j1 + j2 + 1):
                for m1 in range(-j1, j1 + 1):
                    for m2 in range(-j2, j2 + 1):
                        if abs(m1 + m2) <= j3:
                            ref = float(CG(S(j1), S(m1), S(j2), S(m2), S(j3), S(m1 + m2)).doit())
                            worst = max(worst, abs(ref - E.su2_clebsch_gordan_coeff(j1, m1, j2, m2, j3, m1 + m2)))
    assert worst < 1e-14


def test_w3j_identities():
    assert abs(E.wigner_3j(1, 1, 1)[0, 1, 2].item() - 1 / math.sqrt(6)) < 1e-14
    for l in range(5):
        eye = torch.eye(2 * l + 1, dtype=torch.float64) / math.sqrt(2 * l + 1)
        assert (E.wigner_3j(l, 0, l)[:, 0, :] - eye).abs().max() < 1e-14
        assert (E.wigner_3j(0, l, l)[0] - eye).abs().max() < 1e-14
        assert (E.wigner_3j(l, l, 0)[:, :, 0] - eye).abs().max() < 1e-14
    for l1, l2, l3 in [(1, 1, 2), (2, 2, 2), (4, 3, 2), (4, 4, 4)]:
        assert abs(E.wigner_3j(l1, l2, l3).norm().item() - 1) < 1e-14


def test_wigner_D_l1_is_rotation_matrix():
    assert np.abs(E.wigner_D(1, *ANGLES) - E.angles_to_matrix(*ANGLES)).max() < 1e-14


@pytest.mark.parametrize("ls", [(1, 1, 2), (2, 2, 2), (1, 2, 3), (4, 2, 3), (4, 4, 4), (3, 4, 1)])
def test_w3j_equivariance(ls):
    C = E._wigner_3j_np(*ls)
    D = [E.wigner_D(l, *ANGLES) for l in ls]
    assert np.abs(C - np.einsum("ijk,ai,bj,ck->abc", C, *D)).max() < 1e-13


def test_sh_explicit_polynomials_match_recursion_and_are_equivariant():
    torch.manual_seed(0)
    v = torch.randn(64, 3, dtype=torch.float64)
    sh = E.spherical_harmonics(4, v, True, "norm")
    assert (sh - E.spherical_harmonics_recursive(4, v)).abs().max() < 1e-14
    R = torch.from_numpy(E.angles_to_matrix(*ANGLES))
    sh_rot = E.spherical_harmonics(4, v @ R.T, True, "norm")
    off = 0
    for l in range(5):
        D = torch.from_numpy(E.wigner_D(l, *ANGLES))
        blk = slice(off, off + 2 * l + 1)
        assert (sh_rot[:, blk] - sh[:, blk] @ D.T).abs().max() < 1e-13
        assert (sh[:, blk].pow(2).sum(-1) - 1).abs().max() < 1e-13  # 'norm'
        off += 2 * l + 1
    comp = E.spherical_harmonics(4, v, True, "component")
    off = 0
    for l in range(5):
        assert (comp[:, off:off + 2 * l + 1].pow(2).sum(-1) - (2 * l + 1)).abs().max() < 1e-12
        off += 2 * l + 1
    # Y_l(e_y) = e_{m=0}; Y_1 = (x, y, z)
    y = E.spherical_harmonics(4, torch.tensor([[0.0, 1.0, 0.0]], dtype=torch.float64), True, "norm")[0]
    expect = torch.zeros(25, dtype=torch.float64)
    for l in range(5):
        expect[l * l + l] = 1
    assert (y - expect).abs().max() < 1e-14


def test_bessel_basis_known_values():
    x = torch.tensor([1.0, 2.5, 4.999, 5.0, 6.0], dtype=torch.float64)
    out = E.soft_one_hot_linspace_bessel(x, 0.0, 5.0, 8, True)
    n = torch.arange(1, 9, dtype=torch.float64)
    ref = math.sqrt(2 / 5) * torch.sin(n * math.pi * x[:3, None] / 5) / x[:3, None]
    assert (out[:3] - ref).abs().max() < 1e-14
    assert out[3:].abs().max() == 0  # hard cutoff at x >= end


@pytest.mark.parametrize("formula,irreps,dim", [("ijkl=jikl=klij", [(2, 0, 1), (2, 2, 1), (1, 4, 1)], 21),
                                                ("ij=ji", [(1, 0, 1), (1, 2, 1)], 6),
                                                ("ij", [(1, 0, 1), (1, 1, 1), (1, 2, 1)], 9)])
def test_cartesian_tensor_basis(formula, irreps, dim):
    ir, Q = E.reduced_tensor_products(formula)
    assert ir == irreps and Q.shape[0] == dim
    Qf = Q.reshape(dim, -1)
    assert np.abs(Qf @ Qf.T - np.eye(dim)).max() < 1e-12  # orthonormal rows
    rank = Q.ndim - 1
    R = E.angles_to_matrix(*ANGLES)
    Dfull = np.zeros((dim, dim))
    o = 0
    for m, l, _ in ir:
        for _ in range(m):
            Dfull[o:o + 2 * l + 1, o:o + 2 * l + 1] = E.wigner_D(l, *ANGLES)
            o += 2 * l + 1
    letters = "ijkl"[:rank]
    big = "pqrs"[:rank]
    Qr = np.einsum("a" + letters + "," + ",".join(b + c for b, c in zip(big, letters)) + "->a" + big, Q,
                   *([R] * rank))
    assert np.abs(Qr - np.einsum("ba,b...->a...", Dfull, Q)).max() < 1e-12


def test_elasticity_basis_rows_and_symmetry():
    _, Q = E.reduced_tensor_products("ijkl=jikl=klij")
    d = np.eye(3)
    assert np.abs(Q[0] - np.einsum("ij,kl->ijkl", d, d) / 3).max() < 1e-12
    r1 = (np.einsum("ik,jl->ijkl", d, d) + np.einsum("il,jk->ijkl", d, d)) / math.sqrt(20) \
        - np.einsum("ij,kl->ijkl", d, d) / (3 * math.sqrt(5))
    assert np.abs(Q[1] - r1).max() < 1e-12
    for perm in [(0, 2, 1, 3, 4), (0, 1, 2, 4, 3), (0, 3, 4, 1, 2)]:
        assert np.abs(Q - Q.transpose(perm)).max() < 1e-12
    ct = E.CartesianTensor("ijkl=jikl=klij")
    t = torch.randn(5, 21, dtype=torch.float64)
    assert (ct.from_cartesian(ct.to_cartesian(t)) - t).abs().max() < 1e-12


def test_normalize2mom_constants():
    import torch.nn.functional as F

    assert abs(E.normalize2mom(F.silu).cst - 1.679) < 2e-3
    assert abs(E.normalize2mom(torch.tanh).cst - 1.593) < 2e-3
    assert abs(E.normalize2mom(torch.sigmoid).cst - 1.847) < 2e-3
    torch.manual_seed(0)
    z = torch.randn(200000, dtype=torch.float64)
    assert abs(E.normalize2mom(F.silu)(z).pow(2).mean().item() - 1) < 2e-2


def test_tensor_product_component_normalisation_and_equivariance():
    """uvu product of N(0,1) features with component-normalised SH and N(0,1) weights has unit
    second moment per output component (e3nn 'component' + 'element' normalisation)."""
    torch.manual_seed(0)
    in1 = [(8, 0, 1), (8, 1, -1), (4, 2, 1)]
    in2 = [(1, 0, 1), (1, 1, -1), (1, 2, 1)]
    instr, mid = [], []
    for i, (m, l1, p1) in enumerate(in1):
        for j, (_, l2, p2) in enumerate(in2):
            for lo, po in E.irrep_product(l1, p1, l2, p2):
                if lo <= 2:
                    instr.append((i, j, len(mid), "uvu", True))
                    mid.append((m, lo, po))
    tp = E.TensorProduct(in1, in2, mid, instr)
    Z = 20000
    x = torch.randn(Z, E.irreps_dim(in1), dtype=torch.float64)
    v = torch.randn(Z, 3, dtype=torch.float64)
    y = E.spherical_harmonics(2, v, True, "component")
    w = torch.randn(Z, tp.weight_numel, dtype=torch.float64)
    out = tp(x, y, w)
    assert abs(out.pow(2).mean().item() - 1) < 0.05
    # equivariance
    R = torch.from_numpy(E.angles_to_matrix(*ANGLES))

    def D_of(irreps):
        blocks = []
        for m, l, p in irreps:
            blocks += [torch.from_numpy(E.wigner_D(l, *ANGLES))] * m
        return torch.block_diag(*blocks)

    x, y, w = x[:50], y[:50], w[:50]
    y_rot = E.spherical_harmonics(2, v[:50] @ R.T, True, "component")
    out_rot = tp(x @ D_of(in1).T, y_rot, w)
    assert (out_rot - tp(x, y, w) @ D_of(mid).T).abs().max() < 1e-12
