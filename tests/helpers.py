"""Shared fixtures: model configs of the reference, oracle <-> product weight sharing."""
import copy
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# scripts/configs/atomic_tensor.yaml:29-54 backbone (lmax 2) with a pooled rank-2 head
HP_LMAX2 = {
    "species_embedding_dim": 16, "irreps_edge_sh": "0e + 1o + 2e", "num_radial_basis": 8,
    "radial_basis_start": 0.0, "radial_basis_end": 5.0, "radial_basis_type": "bessel", "num_layers": 3,
    "invariant_layers": 2, "invariant_neurons": 32, "average_num_neighbors": 28.0,
    "conv_layer_irreps": "32x0o+32x0e + 16x1o+16x1e + 4x2o+4x2e", "nonlinearity_type": "gate",
    "normalization": "batch", "resnet": True, "conv_to_output_hidden_irreps_out": "16x0e + 2x2e",
    "output_format": "irreps", "output_formula": "ij=ji", "reduce": "mean",
}
# pretrained/20230627/config_final.yaml:24-42 (lmax 4, elasticity tensor)
HP_LMAX4 = {
    "species_embedding_dim": 16, "irreps_edge_sh": "0e + 1o + 2e + 3o + 4e", "num_radial_basis": 8,
    "radial_basis_start": 0.0, "radial_basis_end": 5.0, "radial_basis_type": "bessel", "num_layers": 3,
    "invariant_layers": 2, "invariant_neurons": 32, "average_num_neighbors": 30.4,
    "conv_layer_irreps": "32x0o+32x0e + 16x1o+16x1e + 4x2o+4x2e + 2x3o+2x3e + 2x4e",
    "nonlinearity_type": "gate", "normalization": "batch", "resnet": True,
    "conv_to_output_hidden_irreps_out": "16x0e + 2x2e + 4e", "output_format": "irreps",
    "output_formula": "ijkl=jikl=klij", "reduce": "mean",
}
# tests/model/test_tfn_tensor.py:23-42 of the reference
HP_REFTEST = {
    "species_embedding_dim": 32, "irreps_edge_sh": "0e + 1o + 2e + 3o + 4e", "num_radial_basis": 10,
    "radial_basis_start": 0.0, "radial_basis_end": 5.0, "radial_basis_type": "bessel", "num_layers": 3,
    "invariant_layers": 2, "invariant_neurons": 32, "average_num_neighbors": None,
    "conv_layer_irreps": "32x0o+32x0e+16x1o+16x1e+8x2o+8x2e+4x3o+4x3e+4x4o+4x4e",
    "nonlinearity_type": "gate", "normalization": None, "resnet": True,
    "conv_to_output_hidden_irreps_out": "2x0e + 2x2e + 4e", "output_format": "cartesian",
    "output_formula": "ijkl=jikl=klij", "reduce": "mean",
}
SPECIES8 = [1, 6, 7, 8, 14, 22, 26, 29]


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|  (normwise relative error; the parity metric of DESIGN.md)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.numel() == 0:
        return 0.0
    den = float(b.abs().max())
    return float((a - b).abs().max()) / (den if den > 0 else 1.0)


def elem_err(a: torch.Tensor, b: torch.Tensor, floor: float = 1e-3) -> float:
    """Element-wise error: |a-b| / |b| where |b| > floor * max|b|, |a-b| / (floor * max|b|) below that floor.
    Unlike the normwise metric it does not let small components (high-l blocks, values next to a cut-off) hide
    behind the largest element."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.numel() == 0:
        return 0.0
    scale = float(b.abs().max())
    if scale == 0:
        return float((a - b).abs().max())
    den = b.abs().clamp(min=floor * scale)
    return float(((a - b).abs() / den).max())


def tol(dtype) -> float:
    # BASELINE.json north_star: 1e-5 relative in fp32, 1e-10 in fp64
    return 1e-5 if dtype == torch.float32 else 1e-10


def randomize_bn(model, seed=0):
    """Non-trivial BatchNorm buffers/affine so that eval-mode normalisation is exercised."""
    g = torch.Generator().manual_seed(seed)
    for name, buf in model.named_buffers():
        if name.endswith("running_mean"):
            buf.copy_(torch.randn(buf.shape, generator=g) * 0.3)
        if name.endswith("running_var"):
            buf.copy_(torch.rand(buf.shape, generator=g) + 0.5)
    for name, p in model.named_parameters():
        if ".norm.n.weight" in name:
            p.data.copy_(torch.rand(p.shape, generator=g) + 0.5)
        if ".norm.n.bias" in name:
            p.data.copy_(torch.randn(p.shape, generator=g) * 0.2)


def build_pair(hp, species, dtype, device, seed=0, atomic=False):
    """(oracle model on CPU, product model on `device`) sharing one random state_dict."""
    from matten_b200.model_factory import AtomicTensorModel, ScalarTensorModel
    from oracle import matten_restated as M

    torch.manual_seed(seed)
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        ds = {"allowed_species": species}
        prod = (AtomicTensorModel if atomic else ScalarTensorModel)(hp, ds)
        randomize_bn(prod, seed)
        orac = (M.AtomicTensorModel if atomic else M.ScalarTensorModel)(hp, ds)
    finally:
        torch.set_default_dtype(old)
    missing, unexpected = orac.load_state_dict(prod.state_dict(), strict=False)
    assert not unexpected, unexpected
    # the oracle's e3nn-style modules carry extra (empty / mask) buffers the product does not need
    osd = orac.state_dict()
    assert all(("output_mask" in k) or osd[k].numel() == 0 for k in missing), missing
    prod = prod.to(device=device, dtype=dtype).eval()
    orac = orac.to(dtype=dtype).eval()
    return orac, prod


def load_reference_test_crystal(dtype=torch.float32):
    """tests/test_files/elastic_tensor_one.json of the reference (8-atom TeO3 cell), stored as a
    small fixture under tests/golden (generated by tests/golden/make_golden.py)."""
    with open(os.path.join(GOLDEN, "elastic_tensor_one_structure.json")) as f:
        d = json.load(f)
    from matten_b200.data.neighbors import collate, make_graph

    g = make_graph(np.array(d["cart_coords"]), np.array(d["lattice"]), d["atomic_numbers"], 5.0, dtype=dtype)
    return collate([g]), d


def to_oracle_batch(batch, dtype):
    out = {}
    for k, v in batch.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu()
            out[k] = v.to(dtype) if v.is_floating_point() else v
    return out
