"""The ``torch.library`` registration of the C-ABI entry points (namespace ``matten_b200``): schemas and shape
inference on meta tensors run on CPU; numerics and the registered autograd formulas are checked on the GPU against
the functional layer the modules use."""
import pytest
import torch


def _conv_handle():
    from matten_b200 import o3, ops
    from matten_b200.nn.utils import UVUTensorProduct

    ir = "8x0e+8x1o+4x2e"
    tp = UVUTensorProduct(o3.Irreps(ir), o3.Irreps.spherical_harmonics(2), o3.Irreps(ir), mlp_input_size=8,
                          mlp_hidden_size=16, mlp_num_hidden_layers=2, mlp_activation="silu")
    return tp


def test_ops_are_registered_and_infer_shapes_on_meta():
    from matten_b200 import torch_ops  # noqa: F401  (registers the ops)

    ns = torch.ops.matten_b200
    for name in ("edge_sh", "edge_radial", "csr_by_key", "segment_reduce", "conv_fwd", "conv_bwd", "linear_fwd",
                 "linear_bwd", "gate_fwd"):
        assert hasattr(ns, name), name
    v = torch.empty((7, 3), device="meta")
    assert ns.edge_sh(v, 2, True).shape == (7, 9)
    assert ns.edge_radial(torch.empty(7, device="meta"), 0, 8, 0.0, 5.0, True, 6.0).shape == (7, 8)
    rowptr, perm = ns.csr_by_key(torch.empty(7, dtype=torch.int64, device="meta"), 4)
    assert rowptr.shape == (5,) and perm.shape == (7,) and rowptr.dtype == torch.int32
    assert ns.segment_reduce(torch.empty((6, 5), device="meta"), torch.empty(3, dtype=torch.int32, device="meta"),
                             "mean").shape == (2, 5)


@pytest.mark.gpu
def test_custom_ops_match_functional_layer_and_backpropagate():
    from matten_b200 import ops, torch_ops
    from matten_b200.graph import GraphCache

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    tp = _conv_handle().to(dev)
    N, deg = 40, 9
    E = N * deg
    dst = torch.arange(N).repeat_interleave(deg)
    src = torch.randint(0, N, (E,))
    ei = torch.stack([src, dst]).to(dev)
    x = torch.randn(N, tp.plan.x_dim, device=dev, requires_grad=True)
    sh = ops.edge_sh(torch.randn(E, 3, device=dev), 2, True)
    emb = torch.randn(E, 8, device=dev)
    g = GraphCache({"edge_index": ei, "pos": torch.zeros(N, 3, device=dev)})
    ws = tp.weight_nn.weights()
    pid = torch_ops.register_plan(tp.handle(dev))
    out = torch.ops.matten_b200.conv_fwd(x, sh, emb, ws, g.rowptr, g.perm, g.src_sorted, pid, 9.0, None)
    want = tp.fused(x, sh, emb, g, 9.0)
    assert torch.equal(out, want)
    R = torch.randn_like(out)
    gx, *gw = torch.autograd.grad((out * R).sum(), [x] + ws)
    gx2, *gw2 = torch.autograd.grad((want * R).sum(), [x] + ws)
    assert torch.equal(gx, gx2) and all(torch.equal(a, b) for a, b in zip(gw, gw2))
    # pooling op with its registered backward
    y = torch.randn(12, 5, device=dev, requires_grad=True)
    ptr = torch.tensor([0, 5, 12], dtype=torch.int32, device=dev)
    p = torch.ops.matten_b200.segment_reduce(y, ptr, "mean")
    (gy,) = torch.autograd.grad(p.sum(), y)
    assert torch.allclose(gy[:5], torch.full((5, 5), 0.2, device=dev)) and torch.allclose(gy[5:], torch.full((7, 5), 1 / 7, device=dev))
