"""Backward / training parity on the GPU: gradients of the CUDA path (mt_conv_bwd, mt_linear_bwd, mt_gate_bwd,
training BatchNorm, pooling, species embedding, MSE) against autograd through the CPU oracle with the same
state_dict, and the fused Adam step against torch.optim.Adam.

Tolerances (normwise relative, max|a-b| / max|b| per tensor): 1e-10 in fp64; 2e-5 in fp32 -- gradients are sums over
up to 10^4 edges accumulated in a different (but fixed) order than the oracle's scatter/einsum, so they carry a few
more ulps than the forward's 1e-5.  Whole-model gradients in fp32 pass through four conv layers, three
training-mode BatchNorms and the oracle's own fp32 rounding (the oracle is not exact either): 5e-4 there (mtol)."""
import numpy as np
import pytest
import torch

from tests.helpers import HP_LMAX2, HP_LMAX4, SPECIES8, build_pair, rel_err, to_oracle_batch

pytestmark = pytest.mark.gpu
DTYPES = [torch.float32, torch.float64]


def gtol(dtype):
    return 2e-5 if dtype == torch.float32 else 1e-10


def mtol(dtype):
    return 5e-4 if dtype == torch.float32 else 1e-9


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _conv_pair(dev, dtype, x_ir, sh_lmax, target_ir, S, n_rad, hidden, nlayers, avg, N, deg_fn, seed):
    from matten_b200 import o3
    from matten_b200.nn.conv import PointConv
    from oracle import e3nn_restated as E
    from oracle import matten_restated as M

    torch.manual_seed(seed)
    g = torch.Generator().manual_seed(seed)
    sh_ir = o3.Irreps.spherical_harmonics(sh_lmax)
    irreps_in = {"node_features": o3.Irreps(x_ir), "node_attrs": o3.Irreps(f"{S}x0e"), "edge_attrs": sh_ir,
                 "edge_embedding": o3.Irreps(f"{n_rad}x0e")}
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        conv = PointConv(irreps_in, target_ir, nlayers, hidden, avg)
        oin = {k: o3.parse_irreps_list(v) for k, v in irreps_in.items()}
        ref = M.PointConv(oin, target_ir, nlayers, hidden, avg)
    finally:
        torch.set_default_dtype(old)
    ref.load_state_dict(conv.state_dict(), strict=False)
    dst = torch.cat([torch.full((deg_fn(n),), n, dtype=torch.int64) for n in range(N)])
    E_ = len(dst)
    src = torch.randint(0, N, (E_,), generator=g)
    shuffle = torch.randperm(E_, generator=g)
    ei = torch.stack([src, dst])[:, shuffle]
    x = torch.randn(N, conv.tp.plan.x_dim, dtype=dtype)
    vec = torch.randn(E_, 3, dtype=dtype)
    sh = E.spherical_harmonics(sh_lmax, vec, True, "component")
    emb = torch.randn(E_, n_rad, dtype=dtype)
    sp = torch.randint(0, S, (N,), generator=g)
    num_neigh = torch.bincount(ei[0], minlength=N).to(dtype).clamp(min=1)
    data = {"node_features": x, "node_attrs": torch.nn.functional.one_hot(sp, S).to(dtype), "edge_attrs": sh,
            "edge_embedding": emb, "edge_index": ei, "num_neigh": num_neigh}
    return conv, ref, data, sp


def _conv_grad_case(dev, dtype, *args):
    conv, ref, data, sp = _conv_pair(dev, dtype, *args)
    N = data["node_features"].shape[0]
    # oracle gradients
    xo = data["node_features"].clone().requires_grad_(True)
    out_o = ref(dict(data, node_features=xo))["node_features"]
    R = torch.randn(out_o.shape, dtype=dtype, generator=torch.Generator().manual_seed(99))
    (out_o * R).sum().backward()
    # CUDA path
    conv = conv.to(dev)
    d = {k: v.to(dev) for k, v in data.items()}
    d["species_index"] = sp.to(dev)
    d["pos"] = torch.zeros(N, 3, dtype=dtype, device=dev)
    xg = data["node_features"].to(dev).requires_grad_(True)
    d["node_features"] = xg
    out_g = conv(d)["node_features"]
    assert rel_err(out_g, out_o) < (1e-5 if dtype == torch.float32 else 1e-10)
    (out_g * R.to(dev)).sum().backward()
    assert rel_err(xg.grad, xo.grad) < gtol(dtype), "grad x"
    ref_grads = {k: p.grad for k, p in ref.named_parameters()}
    for k, p in conv.named_parameters():
        assert p.grad is not None, k
        assert rel_err(p.grad, ref_grads[k].reshape(p.grad.shape)) < gtol(dtype), k
    # deterministic: a second backward gives bit-identical gradients
    g1 = {k: p.grad.clone() for k, p in conv.named_parameters()}
    gx1 = xg.grad.clone()
    for p in conv.parameters():
        p.grad = None
    xg.grad = None
    out2 = conv(dict(d, node_features=xg))["node_features"]
    (out2 * R.to(dev)).sum().backward()
    assert torch.equal(gx1, xg.grad)
    for k, p in conv.named_parameters():
        assert torch.equal(g1[k], p.grad), k


@pytest.mark.parametrize("dtype", DTYPES)
def test_conv_backward_lmax2_layers(dev, dtype):
    ir = "32x0o+32x0e+16x1o+16x1e+4x2o+4x2e"
    _conv_grad_case(dev, dtype, "16x0e", 2, "52x0e+16x1o+4x2e", 8, 8, 32, 2, 28.0, 120, lambda n: 28, 0)
    _conv_grad_case(dev, dtype, "32x0e+16x1o+4x2e", 2, "72x0e+16x1o+16x1e+4x2o+4x2e", 8, 8, 32, 2, 28.0, 100,
                    lambda n: 28, 1)
    _conv_grad_case(dev, dtype, ir, 2, ir, 8, 8, 32, 2, 28.0, 90, lambda n: 27 + (n % 3), 2)


@pytest.mark.parametrize("dtype", DTYPES)
def test_conv_backward_lmax4_ragged_and_mlp_variants(dev, dtype):
    ir4 = "32x0o+32x0e+16x1o+16x1e+4x2o+4x2e+2x3o+2x3e+2x4e"
    deg = lambda n: [0, 1, 3, 200, 28, 0, 64, 65][n % 8]  # noqa: E731
    _conv_grad_case(dev, dtype, ir4, 4, ir4, 5, 8, 32, 2, 30.4, 32, deg, 3)
    irt = "32x0o+32x0e+16x1o+16x1e+8x2o+8x2e+4x3o+4x3e+4x4o+4x4e"
    _conv_grad_case(dev, dtype, irt, 4, irt, 2, 10, 32, 2, None, 24, lambda n: 5 + (n % 7), 4)
    ir = "8x0e+8x1o+4x2e"
    _conv_grad_case(dev, dtype, ir, 2, ir, 3, 8, 8, 1, 10.0, 64, lambda n: 12, 5)      # reference default fc sizes
    _conv_grad_case(dev, dtype, ir, 2, ir, 3, 8, 64, 2, 10.0, 64, lambda n: 12, 6)     # 64 hidden units
    _conv_grad_case(dev, dtype, ir, 2, ir, 1, 10, 8, 0, 10.0, 64, lambda n: 12, 7)     # no hidden layer
    _conv_grad_case(dev, dtype, ir, 1, ir, 3, 8, 16, 3, 10.0, 64, lambda n: 12, 8)     # three hidden layers


def _train_pair(hp, species, dtype, dev, seed):
    orac, prod = build_pair(hp, species, dtype, dev, seed=seed)
    return orac.train(), prod.train()


def _model_grads(orac, prod, batch, dtype, dev, key="elastic_tensor_full"):
    from matten_b200 import autograd as A

    ob = to_oracle_batch(batch, dtype)
    db = {k: (v.to(dev).to(dtype) if isinstance(v, torch.Tensor) and v.is_floating_point()
              else (v.to(dev) if isinstance(v, torch.Tensor) else v)) for k, v in batch.items()}
    out_o = orac(ob)
    out_o = out_o[key] if isinstance(out_o, dict) else out_o
    target = torch.randn(out_o.shape, dtype=dtype, generator=torch.Generator().manual_seed(5))
    loss_o = torch.nn.functional.mse_loss(out_o, target)
    loss_o.backward()
    out_g = prod(db)[key]
    loss_g = A.mse_loss(out_g, target.to(dev))
    loss_g.backward()
    return loss_o, loss_g, out_o, out_g


@pytest.mark.parametrize("dtype", DTYPES)
def test_model_backward_lmax2_training_mode(dev, dtype):
    """Whole model in train() mode (BatchNorm batch statistics): loss, every parameter gradient and the updated
    running statistics against the oracle."""
    from matten_b200.data.synthetic import synthetic_batch

    orac, prod = _train_pair(HP_LMAX2, SPECIES8, dtype, dev, seed=3)
    batch = synthetic_batch(4, dtype=dtype)
    loss_o, loss_g, out_o, out_g = _model_grads(orac, prod, batch, dtype, dev)
    assert rel_err(out_g, out_o) < (1e-5 if dtype == torch.float32 else 1e-10)
    assert abs(float(loss_g.detach()) - float(loss_o.detach())) <= gtol(dtype) * abs(float(loss_o.detach()))
    og = {k: p.grad for k, p in orac.named_parameters()}
    n = 0
    for k, p in prod.named_parameters():
        assert p.grad is not None, k
        assert rel_err(p.grad, og[k].reshape(p.grad.shape)) < mtol(dtype), k
        n += 1
    assert n == len(og)
    ob = dict(orac.named_buffers())
    for k, b in prod.named_buffers():
        if k.endswith("running_mean") or k.endswith("running_var"):
            assert rel_err(b, ob[k]) < gtol(dtype), k


@pytest.mark.parametrize("dtype", DTYPES)
def test_model_backward_lmax4_n100_subset(dev, dtype):
    """lmax-4 pretrained architecture on the first 12 example crystals (ragged graphs, many species)."""
    import json
    import os

    from matten_b200.data.neighbors import collate, make_graph
    from tests.helpers import GOLDEN

    with open(os.path.join(GOLDEN, "n100_structures.json")) as f:
        structs = json.load(f)["structures"][:12]
    graphs = [make_graph(np.array(s["cart_coords"]), np.array(s["lattice"]), s["atomic_numbers"], 5.0, dtype)
              for s in structs]
    batch = collate(graphs)
    species = sorted({z for s in structs for z in s["atomic_numbers"]})
    orac, prod = _train_pair(HP_LMAX4, species, dtype, dev, seed=4)
    loss_o, loss_g, out_o, out_g = _model_grads(orac, prod, batch, dtype, dev)
    assert abs(float(loss_g.detach()) - float(loss_o.detach())) <= gtol(dtype) * abs(float(loss_o.detach()))
    og = {k: p.grad for k, p in orac.named_parameters()}
    for k, p in prod.named_parameters():
        assert rel_err(p.grad, og[k].reshape(p.grad.shape)) < mtol(dtype), k


def test_eval_mode_gradients_with_folded_batchnorm(dev):
    """eval() mode under autograd: BatchNorm folded into the gate kernel's affine stage, gradients still exact."""
    from matten_b200.data.synthetic import synthetic_batch

    dtype = torch.float64
    orac, prod = build_pair(HP_LMAX2, SPECIES8, dtype, dev, seed=7)
    batch = synthetic_batch(3, dtype=dtype)
    loss_o, loss_g, _, _ = _model_grads(orac, prod, batch, dtype, dev)
    og = {k: p.grad for k, p in orac.named_parameters()}
    for k, p in prod.named_parameters():
        assert rel_err(p.grad, og[k].reshape(p.grad.shape)) < 1e-9, k


@pytest.mark.parametrize("dtype", DTYPES)
def test_adam_and_mse_kernels(dev, dtype):
    from matten_b200 import ops

    torch.manual_seed(0)
    p0 = torch.randn(10_001, dtype=dtype)
    ref_p = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref_p], lr=0.01, weight_decay=1e-5)
    p = p0.clone().to(dev)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for t in range(1, 6):
        g = torch.randn(10_001, dtype=dtype)
        ref_p.grad = g.clone()
        opt.step()
        ops.adam_step(p, g.to(dev), m, v, t, 0.01, 0.9, 0.999, 1e-8, 1e-5)
    assert rel_err(p, ref_p) < (1e-6 if dtype == torch.float32 else 1e-13)
    a, b = torch.randn(100, 21, dtype=dtype), torch.randn(100, 21, dtype=dtype)
    loss, grad = ops.mse_loss(a.to(dev), b.to(dev))
    ar = a.clone().requires_grad_(True)
    lr_ = torch.nn.functional.mse_loss(ar, b)
    lr_.backward()
    assert abs(float(loss) - float(lr_)) < (1e-6 if dtype == torch.float32 else 1e-13) * float(lr_)
    assert rel_err(grad, ar.grad) < (1e-6 if dtype == torch.float32 else 1e-13)


def test_training_steps_follow_the_oracle(dev):
    """Five optimiser steps (fwd + bwd + Adam, BatchNorm in training mode) of the CUDA trainer against the oracle
    model trained with torch.optim.Adam on the CPU: same loss trajectory and same final parameters (fp64)."""
    from matten_b200.data.synthetic import synthetic_batch
    from matten_b200.train import Trainer

    dtype = torch.float64
    orac, prod = _train_pair(HP_LMAX2, SPECIES8, dtype, dev, seed=11)
    batch = synthetic_batch(4, dtype=dtype)
    ob = to_oracle_batch(batch, dtype)
    db = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
    target = torch.randn(4, 6, dtype=dtype, generator=torch.Generator().manual_seed(1))
    opt = torch.optim.Adam(orac.parameters(), lr=0.01, weight_decay=1e-5)
    tr = Trainer(prod, lr=0.01, weight_decay=1e-5)
    for _ in range(5):
        opt.zero_grad()
        lo = torch.nn.functional.mse_loss(orac(ob), target)
        lo.backward()
        opt.step()
        lg = tr.step(db, target.to(dev))
        assert abs(float(lg) - float(lo)) < 1e-7 * abs(float(lo))
    op = dict(orac.named_parameters())
    for k, p in prod.named_parameters():
        assert rel_err(p, op[k].reshape(p.shape)) < 1e-6, k
