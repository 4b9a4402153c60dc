"""The reference's own tests, run against the oracle restatement (oracle/matten_restated.py):
tests/nn/test_embedding.py:6-13 (integer KAT) and tests/model/test_tfn_tensor.py:98-139
(index symmetry of the predicted elasticity tensor + rotation equivariance, atol 1e-4 in fp32)."""
import numpy as np
import torch

from oracle import e3nn_restated as E
from oracle import matten_restated as M
from tests.helpers import HP_LMAX4, HP_REFTEST, load_reference_test_crystal, to_oracle_batch


def test_atomic_number_to_index_kat():
    n2i = M._AtomicNumberToIndex([6, 1, 8])
    index = n2i(torch.tensor([6, 6, 8, 1, 8]))
    assert index.dtype == torch.long
    assert torch.equal(index, torch.tensor([1, 1, 2, 0, 2]))


def test_appendix_a_path_bookkeeping():
    """uvu paths / weight_numel / D_mid of the reference's test config."""
    model = M.ScalarTensorModel(HP_REFTEST, {"allowed_species": [8, 52]})
    got = []
    for name, mod in model.backbone.named_children():
        pc = getattr(mod, "conv", mod)
        if hasattr(pc, "tp"):
            got.append((len(pc.tp.tp.instructions), pc.tp.tp.weight_numel, E.irreps_dim(pc.tp.irreps_mid)))
    assert got == [(5, 160, 800), (65, 608, 3328), (125, 1056, 5856), (130, 1216, 6656)]
    model = M.ScalarTensorModel(HP_LMAX4, {"allowed_species": [8, 52]})
    got = []
    for name, mod in model.backbone.named_children():
        pc = getattr(mod, "conv", mod)
        if hasattr(pc, "tp"):
            got.append((len(pc.tp.tp.instructions), pc.tp.tp.weight_numel, E.irreps_dim(pc.tp.irreps_mid)))
    assert got == [(5, 80, 400), (59, 452, 2324), (99, 714, 3658), (103, 842, 4170)]


def test_model_equivariance_and_symmetry_fp32():
    torch.manual_seed(35)
    model = M.ScalarTensorModel(HP_REFTEST, {"allowed_species": [8, 52]}).eval()
    batch, raw = load_reference_test_crystal(torch.float32)
    b = to_oracle_batch(batch, torch.float32)
    Q = torch.tensor(E.angles_to_matrix(0.3, 1.1, -0.7), dtype=torch.float32)
    b_rot = dict(b)
    b_rot["pos"] = b["pos"] @ Q.T
    b_rot["cell"] = b["cell"] @ Q.T
    with torch.no_grad():
        pred = model(b)[0]
        pred_rot = model(b_rot)[0]
    assert torch.allclose(pred, pred.swapaxes(0, 1))
    assert torch.allclose(pred, pred.swapaxes(2, 3))
    assert torch.allclose(pred, pred.swapaxes(0, 2).swapaxes(1, 3))
    x = torch.einsum("im,jn,kp,lq,mnpq->ijkl", Q, Q, Q, Q, pred)
    assert torch.allclose(x, pred_rot, atol=1e-4)
    assert pred.abs().max() > 1e-3  # not trivially zero


def test_golden_regression():
    """oracle fp64 output pinned in tests/golden/oracle_lmax2_seed0.pt (made by make_golden.py)."""
    import os

    from tests.helpers import GOLDEN, HP_LMAX2, SPECIES8

    g = torch.load(os.path.join(GOLDEN, "oracle_lmax2_seed0.pt"), weights_only=False)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        model = M.ScalarTensorModel(HP_LMAX2, {"allowed_species": SPECIES8})
    finally:
        torch.set_default_dtype(old)
    model.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in g["state_dict"].items()},
                          strict=False)
    model.eval()
    with torch.no_grad():
        out = model({k: v for k, v in g["batch"].items() if isinstance(v, torch.Tensor)})
    assert (out - g["output"]).abs().max() < 1e-12
