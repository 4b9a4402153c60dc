"""Pinning the oracle restatements of the SURVEY section 8(f4) rows on the CPU by their defining properties and
hand-computed values (the reference has no tests of its own for these layers; e3nn / torch_geometric / sklearn are
not installed here):

* graph InstanceNorm (reference src/matten/nn/utils.py:448-588): hand-computed two-graph example; per-graph channel
  statistics after normalisation; independence of one graph from the others; rotation equivariance;
* NormActivation (:142-150): direction is kept, the norm is mapped through f; equivariance; the epsilon clamp;
* MeanNormNormalize / ScalarNormalize (src/matten/data/transform.py:59-302): closed-form statistics, round trip."""
import math

import numpy as np
import pytest
import torch

from oracle import e3nn_restated as E
from oracle import matten_restated as M


def _rot_blocks(irreps, a, b, c):
    mats = []
    for m, l, p in E.parse_irreps(irreps):
        D = torch.from_numpy(E.wigner_D(l, a, b, c))
        mats += [D] * m
    return torch.block_diag(*mats)


def test_instance_norm_hand_example():
    # graph 0: nodes (1, 3) on the 0e channel, vectors (3,0,0), (0,4,0); graph 1: a single node
    x = torch.tensor([[1.0, 3, 0, 0], [3.0, 0, 4, 0], [5.0, 1, 2, 2]], dtype=torch.float64)
    batch = torch.tensor([0, 0, 1])
    n = M.InstanceNorm("1x0e+1x1o", eps=0.0, affine=False)
    y = n(x, batch)
    # scalars: mean 2, centred (-1, 1), mean square 1 -> (-1, 1).  vectors: component-mean squares 3 and 16/3,
    # graph mean 25/6 -> scale sqrt(6)/5
    s = math.sqrt(6) / 5
    want01 = torch.tensor([[-1.0, 3 * s, 0, 0], [1.0, 0, 4 * s, 0]], dtype=torch.float64)
    assert torch.allclose(y[:2], want01, atol=1e-14)
    assert torch.allclose(y[2, 1:], x[2, 1:] / math.sqrt(3.0), atol=1e-14)  # |v|^2/3 = 3
    n2 = M.InstanceNorm("1x0e+1x1o", eps=0.0, affine=False, reduce="max", normalization="norm")
    y2 = n2(x, batch)
    assert torch.allclose(y2[:2, 1:], x[:2, 1:] / 4.0, atol=1e-14)  # max |v|^2 = 16


@pytest.mark.parametrize("reduce", ["mean", "max"])
@pytest.mark.parametrize("normalization", ["component", "norm"])
def test_instance_norm_statistics_locality_equivariance(reduce, normalization):
    torch.manual_seed(0)
    irreps = "3x0e+2x0o+2x1o+1x2e"
    dim = E.irreps_dim(E.parse_irreps(irreps))
    batch = torch.tensor([0] * 5 + [1] * 9 + [2] * 3)
    x = torch.randn(len(batch), dim, dtype=torch.float64) + 0.5
    n = M.InstanceNorm(irreps, eps=0.0, affine=False, reduce=reduce, normalization=normalization)
    y = n(x, batch)
    for g in range(3):
        yg = y[batch == g]
        assert torch.allclose(yg[:, :5].mean(0), torch.zeros(5, dtype=torch.float64), atol=1e-13)  # l = 0, both parities
        col = 0
        for m, l, p in E.parse_irreps(irreps):
            d = 2 * l + 1
            f = yg[:, col:col + m * d].reshape(-1, m, d)
            col += m * d
            sq = f.pow(2).sum(-1) if normalization == "norm" else f.pow(2).mean(-1)
            stat = sq.mean(0) if reduce == "mean" else sq.max(0).values
            assert torch.allclose(stat, torch.ones(m, dtype=torch.float64), atol=1e-12)
    # a graph's output does not depend on the other graphs
    x2 = x.clone()
    x2[batch == 1] = torch.randn(9, dim, dtype=torch.float64)
    y2 = n(x2, batch)
    assert torch.equal(y2[batch != 1], y[batch != 1])
    # rotation equivariance (with affine parameters)
    na = M.InstanceNorm(irreps, reduce=reduce, normalization=normalization).double()
    na.weight.data.uniform_(0.5, 1.5)
    na.bias.data.normal_()
    R = _rot_blocks(irreps, 0.3, 1.1, -0.7)
    assert torch.allclose(na(x @ R.T, batch), na(x, batch) @ R.T, atol=1e-12)
    assert na.bias.shape == (5,) and na.weight.shape == (8,)


def test_norm_activation_properties():
    torch.manual_seed(1)
    irreps = "4x0e+3x1o+2x2e"
    ir = E.parse_irreps(irreps)
    x = torch.randn(50, E.irreps_dim(ir), dtype=torch.float64)
    for f in (M.ACTIVATION["e"]["ssp"], M.ACTIVATION["e"]["silu"], M.ACTIVATION["e"]["sigmoid"]):
        na = E.NormActivation(ir, f, normalize=True, epsilon=1e-8, bias=False)
        y = na(x)
        col = 0
        for m, l, p in ir:
            d = 2 * l + 1
            xi = x[:, col:col + m * d].reshape(-1, m, d)
            yi = y[:, col:col + m * d].reshape(-1, m, d)
            col += m * d
            nx = xi.norm(dim=-1)
            assert torch.allclose(yi.norm(dim=-1), f(nx).abs(), atol=1e-12)          # |y| = |f(|x|)|
            assert torch.allclose(yi * nx[..., None], xi * f(nx)[..., None], atol=1e-12)  # same direction
        R = _rot_blocks(irreps, -0.4, 0.8, 2.0)
        assert torch.allclose(na(x @ R.T), y @ R.T, atol=1e-12)
    # scalars: sign(x) f(|x|); a zero row stays zero with the clamp (no 0/0)
    na = E.NormActivation("2x0e", torch.sigmoid)
    v = torch.tensor([[2.0, -2.0], [0.0, 0.0]], dtype=torch.float64)
    out = na(v)
    s2 = float(torch.sigmoid(torch.tensor(2.0, dtype=torch.float64)))
    assert torch.allclose(out, torch.tensor([[s2, -s2], [0.0, 0.0]], dtype=torch.float64))
    with pytest.raises(ValueError):
        E.NormActivation("1x0e", torch.sigmoid, normalize=False, epsilon=1e-8)
    # ActivationLayer(activation_type="norm"): scalars + gated irreps the product can reach, even-scalar activation
    act = M.ActivationLayer("8x0e+4x1o", "0e+1o", "8x0e+8x0o+4x1o+4x1e+2x2e", activation_type="norm",
                            activation_scalars={"e": "silu", "o": "tanh"})
    assert E.irreps_str(act.irreps_in) == "8x0e+4x1o+4x1e+2x2e" == E.irreps_str(act.irreps_out)
    with pytest.raises(ValueError, match="Support `activation_type`"):
        M.ActivationLayer("8x0e", "0e", "8x0e", activation_type="relu")


def test_target_normalizers_closed_form_and_round_trip():
    torch.manual_seed(2)
    irreps = "2x0e+2x2e+4e"
    data = torch.randn(400, 21, dtype=torch.float64) * 3 + 1
    n = M.MeanNormNormalize(irreps, eps=0.0)
    with pytest.raises(RuntimeError, match="mean and norm not initialized"):
        n(data)
    mean, norm = n.compute_statistics(data)
    assert torch.allclose(mean[:2], data[:, :2].mean(0)) and float(mean[2:].abs().max()) == 0
    assert torch.allclose(norm[:2], data[:, :2].std(0, unbiased=False))
    for lo, hi in ((2, 7), (7, 12), (12, 21)):  # one norm per irrep copy: root mean square over samples and components
        assert torch.allclose(norm[lo:hi], data[:, lo:hi].pow(2).mean().sqrt().expand(hi - lo))
    y = n(data)
    assert torch.allclose(y[:, :2].mean(0), torch.zeros(2, dtype=torch.float64), atol=1e-13)
    assert torch.allclose(n.inverse(y), data, atol=1e-12)
    nn_ = M.MeanNormNormalize(irreps, eps=0.0, normalization="norm")
    _, norm2 = nn_.compute_statistics(data)
    assert torch.allclose(norm2[2:7], norm[2:7] * math.sqrt(5)) and torch.allclose(norm2[12:], norm[12:] * 3)
    s = M.ScalarNormalize(3)
    d3 = torch.stack([torch.arange(5.0), torch.full((5,), 2.0), torch.tensor([1.0, -1, 1, -1, 0])], 1)
    mean, std = s.compute_statistics(d3)
    assert np.allclose(mean.numpy(), [2.0, 2.0, 0.0]) and np.allclose(std.numpy(), [math.sqrt(2.0), 1.0, math.sqrt(0.8)])
    assert torch.allclose(s.inverse(s(d3)), d3, atol=1e-6)  # fp32 data
