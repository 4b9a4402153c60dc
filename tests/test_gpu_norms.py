"""GPU parity for the SURVEY section 8(f4) rows: graph InstanceNorm (reference src/matten/nn/utils.py:448-588),
NormActivation (:142-150), the target normalisers (src/matten/data/transform.py:59-302) and the arg-based backward of
min/max pooling -- forward and backward through the C ABI against the CPU oracle (fp64 autograd) on the same seeded
inputs.  Tolerances: BASELINE.json's 1e-5 (fp32) / 1e-10 (fp64) normwise, element-wise where stated; the
elementwise normaliser maps are bit exact."""
import copy

import pytest
import torch

from tests.helpers import HP_LMAX2, SPECIES8, build_pair, elem_err, rel_err, to_oracle_batch, tol

pytestmark = pytest.mark.gpu
DTYPES = [torch.float32, torch.float64]
DEV = "cuda:0"


def _ragged_batch(counts):
    return torch.cat([torch.full((c,), i, dtype=torch.int64) for i, c in enumerate(counts)])


# ------------------------------------------------------------------ InstanceNorm
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("reduce", ["mean", "max"])
@pytest.mark.parametrize("normalization", ["component", "norm"])
@pytest.mark.parametrize("affine", [True, False])
def test_instance_norm_forward_backward(dtype, reduce, normalization, affine):
    from matten_b200.nn.utils import InstanceNorm
    from oracle import matten_restated as M

    torch.manual_seed(11)
    irreps = "5x0e+3x0o+4x1o+2x2e+1x3o+200x0e+3x4e"
    batch = _ragged_batch([7, 1, 33, 2, 64, 5])
    mod = InstanceNorm(irreps, reduce=reduce, normalization=normalization, affine=affine).to(dtype)
    ref = M.InstanceNorm(irreps, reduce=reduce, normalization=normalization, affine=affine).double()
    if affine:
        mod.weight.data.uniform_(0.5, 1.5)
        mod.bias.data.normal_()
        ref.load_state_dict({k: v.double() for k, v in mod.state_dict().items()})
        assert mod.bias.shape == (208,) and mod.weight.shape == (5 + 3 + 4 + 2 + 1 + 200 + 3,)
    x = (torch.randn(len(batch), mod.irreps.dim, dtype=dtype) * 2 + 0.3)
    go = torch.randn_like(x)
    xr = x.detach().clone().double().requires_grad_(True)
    want = ref(xr, batch)
    want.backward(go.double())
    mod = mod.to(DEV)
    xd = x.to(DEV).requires_grad_(True)
    got = mod(xd, batch.to(DEV))
    got.backward(go.to(DEV))
    assert rel_err(got, want) < tol(dtype)
    assert elem_err(got, want) < 20 * tol(dtype)
    assert rel_err(xd.grad, xr.grad) < 4 * tol(dtype)
    if affine:
        assert rel_err(mod.weight.grad, ref.weight.grad) < 4 * tol(dtype)
        assert rel_err(mod.bias.grad, ref.bias.grad) < 4 * tol(dtype)
    # inference path (no autograd) gives the same values; training flag changes nothing (reference note :440-441)
    with torch.no_grad():
        assert torch.equal(mod.eval()(x.to(DEV), batch.to(DEV)), got.detach())


def test_instance_norm_unsorted_batch_is_rejected():
    from matten_b200.nn.utils import InstanceNorm

    mod = InstanceNorm("2x0e+1x1o").to(DEV)
    x = torch.randn(4, 5, device=DEV)
    with pytest.raises(Exception):
        mod(x, torch.tensor([0, 1, 0, 1], device=DEV))
    with pytest.raises(AssertionError):
        mod(torch.randn(4, 6, device=DEV), torch.tensor([0, 0, 1, 1], device=DEV))


# ------------------------------------------------------------------ NormActivation
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("act", ["ssp", "silu", "sigmoid"])
def test_norm_activation_forward_backward(dtype, act):
    from matten_b200.nn.utils import ActivationLayer
    from oracle import matten_restated as M

    torch.manual_seed(5)
    x_ir, sh, out_ir = "8x0o+8x0e+4x1o+4x1e+2x2e", "0e+1o+2e", "8x0o+8x0e+4x1o+4x1e+2x2o+2x2e+2x3o"
    kw = dict(activation_type="norm", activation_scalars={"e": act, "o": "tanh"})
    mod = ActivationLayer(x_ir, sh, out_ir, **kw)
    ref = M.ActivationLayer(x_ir, sh, out_ir, **kw)
    assert str(mod.irreps_in) == str(mod.irreps_out)
    assert mod.irreps_in.dim == sum(m * (2 * l + 1) for m, l, _ in ref.irreps_in)
    x = torch.randn(257, mod.irreps_in.dim, dtype=dtype) * 1.5
    x[3] = 0                       # zero norm: clamped at epsilon, gradient is f(eps)/eps * g
    x[4] *= 1e-12                  # below the clamp
    x[5] *= 30                     # softplus threshold region
    go = torch.randn_like(x)
    xr = x.detach().clone().double().requires_grad_(True)
    want = ref(xr)
    want.backward(go.double())
    xd = x.to(DEV).requires_grad_(True)
    got = mod.to(DEV)(xd)
    got.backward(go.to(DEV))
    assert rel_err(got, want) < tol(dtype)
    assert elem_err(got, want) < 20 * tol(dtype)
    keep = torch.ones(len(x), dtype=torch.bool)
    keep[3:5] = False
    assert rel_err(xd.grad.cpu()[keep], xr.grad[keep]) < 4 * tol(dtype)
    # clamped rows: the scale is f(1e-8) / 1e-8; softplus(1e-8) - log 2 in the oracle cancels to ~1e-8 relative
    # even in fp64 (the kernel evaluates log1p(expm1(n) / 2) instead), so these two rows get that bound
    assert rel_err(xd.grad.cpu()[~keep], xr.grad[~keep]) < 1e-6
    with torch.no_grad():
        assert torch.equal(mod(x.to(DEV)), got.detach())


@pytest.mark.parametrize("dtype", DTYPES)
def test_model_with_norm_activation_and_instance_norm(dtype):
    """The two optional layer kinds inside the whole model (create_model options ``nonlinearity_type: norm`` and
    ``normalization: instance``, reference src/matten/model_factory/tfn_scalar_tensor.py:120-135)."""
    from matten_b200.data.synthetic import synthetic_batch

    hp = copy.deepcopy(HP_LMAX2)
    hp["nonlinearity_type"], hp["normalization"] = "norm", "instance"
    orac, prod = build_pair(hp, SPECIES8, dtype, torch.device(DEV), seed=2)
    batch = synthetic_batch(5, dtype=dtype)
    ob = to_oracle_batch(batch, dtype)
    db = {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
    with torch.no_grad():
        want = orac(ob)
        got = prod(db)["elastic_tensor_full"]
    assert rel_err(got, want) < 2 * tol(dtype)


# ------------------------------------------------------------------ target normalisers
@pytest.mark.parametrize("dtype", DTYPES)
def test_mean_norm_normalize(dtype):
    from matten_b200.data.transform import MeanNormNormalize, TensorTargetTransform
    from oracle import matten_restated as M

    torch.manual_seed(2)
    irreps = "2x0e+2x2e+4e"  # the elasticity target (reference transform.py:544)
    data = torch.randn(1000, 21, dtype=dtype) * torch.linspace(0.5, 40, 21, dtype=dtype) + 7
    for normalization in ("component", "norm"):
        ref = M.MeanNormNormalize(irreps, normalization=normalization, scale=0.7)
        mean_r, norm_r = ref.compute_statistics(data.double())
        mod = MeanNormNormalize(irreps, normalization=normalization, scale=0.7)
        with pytest.raises(RuntimeError, match="mean and norm not initialized"):
            mod(data.to(DEV))
        mean, norm = mod.compute_statistics(data.to(DEV))
        assert rel_err(mean, mean_r) < tol(dtype) and rel_err(norm, norm_r) < tol(dtype)
        assert float(mean[2:].abs().max()) == 0.0  # only the 0e channels are centred
        # elementwise maps: bit exact against the same IEEE expression on the host
        m_h, n_h = mean.cpu(), norm.cpu()
        fwd = mod(data.to(DEV))
        assert torch.equal(fwd.cpu(), (data - m_h) / (n_h * 0.7))
        inv = mod.inverse(fwd)
        assert torch.equal(inv.cpu(), fwd.cpu() * (n_h * 0.7) + m_h)
        assert rel_err(inv, data) < 10 * torch.finfo(dtype).eps * 50
    with pytest.raises(ValueError, match="Invalid reduce option"):
        MeanNormNormalize(irreps, reduce="max").compute_statistics(data.to(DEV))
    # state-dict round trip into the target transform (dataset_statistics.pt layout)
    t = TensorTargetTransform(irreps=irreps)
    with pytest.raises(ValueError, match="Cannot load dataset statistics"):
        t(data.to(DEV))
    stats = t.compute_statistics(data.to(DEV), atomic_numbers=[14, 8, 8], num_neigh=[30.0, 28.0])
    assert set(stats) == {"elastic_tensor_full", "allowed_species", "average_num_neigh"}
    t2 = TensorTargetTransform(irreps=irreps)
    t2.normalizer.load_state_dict(stats["elastic_tensor_full"])
    assert torch.equal(t2.inverse(t2(data.to(DEV))), t.inverse(t(data.to(DEV))))


@pytest.mark.parametrize("dtype", DTYPES)
def test_scalar_normalize(dtype):
    from matten_b200.data.transform import ScalarNormalize, ScalarTargetTransform
    from oracle import matten_restated as M

    torch.manual_seed(3)
    data = torch.randn(777, 4, dtype=dtype) * torch.tensor([1.0, 20.0, 0.01, 0.0], dtype=dtype) + 3
    ref = M.ScalarNormalize(4)
    mean_r, std_r = ref.compute_statistics(data)
    mod = ScalarNormalize(4)
    mean, std = mod.compute_statistics(data.to(DEV))
    assert rel_err(mean, mean_r) < tol(dtype) and rel_err(std, std_r) < tol(dtype)
    assert float(std[3]) == 1.0  # constant feature: deviation replaced by 1 (sklearn StandardScaler)
    fwd = mod(data.to(DEV))
    assert torch.equal(fwd.cpu(), (data - mean.cpu()) / (std.cpu() * 1.0))
    tt = ScalarTargetTransform(["a", "b"])
    tt.compute_statistics({"a": data[:, :1].to(DEV), "b": data[:, 1:2].to(DEV)})
    out = tt({"a": data[:, :1].to(DEV), "b": data[:, 1:2].to(DEV)})
    assert rel_err(tt.inverse(out["b"], "b"), data[:, 1:2]) < 1e-5


# ------------------------------------------------------------------ min / max pooling backward
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("mode", ["min", "max"])
def test_min_max_pooling_backward(dtype, mode):
    from matten_b200 import functional as F
    from matten_b200 import ops
    from oracle import e3nn_restated as E

    torch.manual_seed(4)
    counts = [3, 1, 0, 64, 7, 0]
    batch = _ragged_batch(counts)
    x = torch.randn(len(batch), 35, dtype=dtype)
    go = torch.randn(len(counts), 35, dtype=dtype)
    xr = x.clone().requires_grad_(True)
    E.scatter(xr, batch, dim_size=len(counts), reduce=mode).backward(go)
    ptr, _ = ops.csr_by_key(batch.to(DEV), len(counts), False, ops.new_flag(DEV))
    xd = x.to(DEV).requires_grad_(True)
    F.segment_reduce(xd, ptr, mode).backward(go.to(DEV))
    assert torch.equal(xd.grad.cpu(), xr.grad)  # routing of the gradient: bit exact


# ------------------------------------------------------------------ trainable Bessel frequencies
@pytest.mark.parametrize("dtype", DTYPES)
def test_trainable_bessel_frequency_gradient(dtype):
    """BesselBasis(trainable=True) (reference src/matten/nn/_nequip.py:80-126): gradient of the edge embedding with
    respect to the frequencies, against fp64 autograd of the oracle."""
    from matten_b200.nn._nequip import RadialBasisEdgeEncoding
    from oracle import matten_restated as M

    torch.manual_seed(6)
    E_ = 4097
    r = (torch.rand(E_, dtype=dtype) * 5.5 + 0.3)  # some edges beyond the cut-off
    mod = RadialBasisEdgeEncoding(basis_kwargs={"r_max": 5.0, "num_basis": 8, "trainable": True},
                                  cutoff_kwargs={"r_max": 5.0, "p": 6}).to(DEV)
    mod.basis.to(dtype)
    ref = M.RadialBasisEdgeEncoding({"r_max": 5.0, "num_basis": 8, "trainable": True}, {"r_max": 5.0, "p": 6}).double()
    with torch.no_grad():
        mod.basis.bessel_weights += torch.linspace(-0.2, 0.3, 8, device=DEV)
        ref.basis.bessel_weights.copy_(mod.basis.bessel_weights.double().cpu())
    go = torch.randn(E_, 8, dtype=dtype)
    vec = torch.zeros(E_, 3, dtype=dtype)
    vec[:, 0] = r
    want = ref({"edge_vectors": vec.double(), "edge_lengths": r.double()})["edge_embedding"]
    want.backward(go.double())
    got = mod({"edge_vectors": vec.to(DEV), "edge_lengths": r.to(DEV)})["edge_embedding"]
    got.backward(go.to(DEV))
    assert rel_err(got, want) < tol(dtype)
    assert rel_err(mod.basis.bessel_weights.grad, ref.basis.bessel_weights.grad) < 4 * tol(dtype)
