"""End-to-end GPU parity: the product model (CUDA kernels behind the reference's module API)
against the CPU oracle with the same state_dict, plus the reference's own property tests
(tests/model/test_tfn_tensor.py:98-139) run on the CUDA path, determinism and golden regression."""
import os

import numpy as np
import pytest
import torch

from tests.helpers import (GOLDEN, HP_LMAX2, HP_LMAX4, HP_REFTEST, SPECIES8, build_pair,
                           load_reference_test_crystal, rel_err, to_oracle_batch, tol)

pytestmark = pytest.mark.gpu
DTYPES = [torch.float32, torch.float64]


def _dev_batch(batch, dev, dtype):
    return {k: (v.to(dev).to(dtype) if isinstance(v, torch.Tensor) and v.is_floating_point()
                else (v.to(dev) if isinstance(v, torch.Tensor) else v)) for k, v in batch.items()}


def _layerwise(orac, prod, ob, db):
    """node features after every backbone module."""
    feats = []
    o, d = dict(ob), dict(db)
    for (n1, m1), (n2, m2) in zip(orac.backbone.named_children(), prod.backbone.named_children()):
        o = m1(o)
        d = m2(d)
        feats.append((n1, o.get("node_features"), d.get("node_features")))
    return feats, o, d


@pytest.mark.parametrize("dtype", DTYPES)
def test_lmax2_model_parity_synthetic(dtype):
    from matten_b200.data.synthetic import synthetic_batch

    dev = torch.device("cuda:0")
    orac, prod = build_pair(HP_LMAX2, SPECIES8, dtype, dev)
    batch = synthetic_batch(6, dtype=dtype)
    ob, db = to_oracle_batch(batch, dtype), _dev_batch(batch, dev, dtype)
    with torch.no_grad():
        feats, _, _ = _layerwise(orac, prod, ob, db)
        for name, a, b in feats:
            assert rel_err(b, a) < tol(dtype), name
        want = orac(ob)
        got = prod(db)["elastic_tensor_full"]
    assert got.shape == (6, 6)
    assert rel_err(got, want) < tol(dtype)
    # edge/index bookkeeping is bit exact
    d2 = dict(db)
    with torch.no_grad():
        prod.backbone(d2)
    g = d2["_mt_graph"]
    dst = batch["edge_index"][1].numpy()
    perm = np.argsort(dst, kind="stable")
    assert np.array_equal(g.perm.cpu().numpy(), perm.astype(np.int32))
    assert np.array_equal(g.src_sorted.cpu().numpy(), batch["edge_index"][0].numpy()[perm].astype(np.int32))
    assert np.array_equal(g.rowptr.cpu().numpy(),
                          np.concatenate([[0], np.cumsum(np.bincount(dst, minlength=len(batch["pos"])))]))


@pytest.mark.parametrize("dtype", DTYPES)
def test_lmax4_elasticity_model_parity_n100(dtype):
    """BASELINE config 1 workload: the 100 example crystals (473 atoms, 73 elements), lmax-4
    architecture of the pretrained model, random weights (the checkpoint is not available offline)."""
    import json

    from matten_b200.data.neighbors import collate, make_graph
    from matten_b200.nn.readout import CartesianTensorWrapper

    dev = torch.device("cuda:0")
    with open(os.path.join(GOLDEN, "n100_structures.json")) as f:
        structs = json.load(f)["structures"]
    graphs = [make_graph(np.array(s["cart_coords"]), np.array(s["lattice"]), s["atomic_numbers"], 5.0, dtype)
              for s in structs]
    batch = collate(graphs)
    assert batch["pos"].shape[0] == 473 and batch["edge_index"].shape[1] == 14380
    species = sorted({z for s in structs for z in s["atomic_numbers"]})
    orac, prod = build_pair(HP_LMAX4, species, dtype, dev, seed=1)
    ob, db = to_oracle_batch(batch, dtype), _dev_batch(batch, dev, dtype)
    with torch.no_grad():
        feats, _, _ = _layerwise(orac, prod, ob, db)
        for name, a, b in feats:
            assert rel_err(b, a) < tol(dtype), name
        want = orac(ob)
        got = prod(db)["elastic_tensor_full"]
        assert got.shape == (100, 21)
        assert rel_err(got, want) < tol(dtype)
        cart = CartesianTensorWrapper("ijkl=jikl=klij").to_cartesian(got)
        assert rel_err(cart, orac.ct.to_cartesian(want)) < tol(dtype)
        assert cart.shape == (100, 3, 3, 3, 3)


def test_reference_property_test_on_cuda_path():
    """tests/model/test_tfn_tensor.py of the reference: lmax-4 test config, no normalisation,
    per-node num_neigh; symmetric output and f(Qx) = Q f(x) with atol 1e-4 (fp32)."""
    from oracle import e3nn_restated as E

    dev = torch.device("cuda:0")
    torch.manual_seed(35)
    orac, prod = build_pair(HP_REFTEST, [8, 52], torch.float32, dev, seed=35)
    batch, _ = load_reference_test_crystal(torch.float32)
    Q = torch.tensor(E.angles_to_matrix(0.3, 1.1, -0.7), dtype=torch.float32)
    rot = dict(batch)
    rot["pos"] = batch["pos"] @ Q.T
    rot["cell"] = batch["cell"] @ Q.T
    with torch.no_grad():
        pred = prod(_dev_batch(batch, dev, torch.float32))["elastic_tensor_full"][0].cpu()
        pred_rot = prod(_dev_batch(rot, dev, torch.float32))["elastic_tensor_full"][0].cpu()
        want = orac(to_oracle_batch(batch, torch.float32))[0]
    assert pred.shape == (3, 3, 3, 3)
    assert rel_err(pred, want) < 1e-5
    assert torch.allclose(pred, pred.swapaxes(0, 1), atol=1e-7)
    assert torch.allclose(pred, pred.swapaxes(2, 3), atol=1e-7)
    assert torch.allclose(pred, pred.swapaxes(0, 2).swapaxes(1, 3), atol=1e-7)
    x = torch.einsum("im,jn,kp,lq,mnpq->ijkl", Q, Q, Q, Q, pred)
    assert torch.allclose(x, pred_rot, atol=1e-4)


def test_atomic_tensor_model_parity():
    """BASELINE config 4: per-atom rank-2 outputs (scripts/configs/atomic_tensor.yaml)."""
    from matten_b200.data.synthetic import synthetic_batch

    dev = torch.device("cuda:0")
    hp = dict(HP_LMAX2, output_format="cartesian")
    orac, prod = build_pair(hp, [8, 14], torch.float32, dev, seed=2, atomic=True)
    batch = synthetic_batch(3, species=[8, 14], dtype=torch.float32)
    with torch.no_grad():
        want = orac(to_oracle_batch(batch, torch.float32))
        got = prod(_dev_batch(batch, dev, torch.float32))["nmr_tensor"]
    assert got.shape == (192, 3, 3)
    assert rel_err(got, want) < 1e-5
    assert torch.allclose(got, got.transpose(1, 2), atol=1e-6)


def test_determinism_and_bad_species():
    from matten_b200.data.synthetic import synthetic_batch

    dev = torch.device("cuda:0")
    _, prod = build_pair(HP_LMAX2, SPECIES8, torch.float32, dev)
    batch = synthetic_batch(4)
    db = _dev_batch(batch, dev, torch.float32)
    with torch.no_grad():
        a = prod(db)["elastic_tensor_full"]
        b = prod(db)["elastic_tensor_full"]
    assert torch.equal(a, b)
    bad = dict(db)
    bad["atomic_numbers"] = db["atomic_numbers"].clone()
    bad["atomic_numbers"][3] = 3  # Li is not in the species list
    with pytest.raises(RuntimeError, match="atomic numbers"):
        with torch.no_grad():
            prod(bad)


def test_golden_fixture_fp64():
    """the committed oracle fixture (tests/golden/make_golden.py) through the CUDA path, fp64."""
    from matten_b200.model_factory import ScalarTensorModel

    dev = torch.device("cuda:0")
    g = torch.load(os.path.join(GOLDEN, "oracle_lmax2_seed0.pt"), weights_only=False)
    model = ScalarTensorModel(HP_LMAX2, {"allowed_species": SPECIES8})
    missing, unexpected = model.load_state_dict(g["state_dict"], strict=False)
    assert not missing, missing
    model = model.to(dev).double().eval()
    with torch.no_grad():
        out = model(_dev_batch(g["batch"], dev, torch.float64))["elastic_tensor_full"]
    assert rel_err(out, g["output"]) < 1e-10


def test_full_size_properties():
    """BASELINE config 2 at full size (512 x 64 atoms, 917 504 edges): size-independent checks --
    permuting the crystals permutes the outputs; a crystal's output does not depend on its batch."""
    from matten_b200.data.synthetic import synthetic_batch, tile_batch

    dev = torch.device("cuda:0")
    _, prod = build_pair(HP_LMAX2, SPECIES8, torch.float32, dev)
    small = synthetic_batch(8)
    big = tile_batch(small, 64)
    assert big["pos"].shape[0] == 32768 and big["edge_index"].shape[1] == 917504
    with torch.no_grad():
        out_small = prod(_dev_batch(small, dev, torch.float32))["elastic_tensor_full"]
        out_big = prod(_dev_batch(big, dev, torch.float32))["elastic_tensor_full"]
    assert out_big.shape == (512, 6)
    assert torch.isfinite(out_big).all()
    # every tile of 8 crystals is a copy of the small batch: identical results, bit for bit
    assert torch.equal(out_big.reshape(64, 8, 6), out_small.unsqueeze(0).expand(64, 8, 6))


def test_cuda_graph_replay_is_bitwise_equal_and_keeps_the_error_word():
    """``CapturedForward`` (what bench.py times): the replayed graph gives bit-identical predictions to the
    kernel-by-kernel forward, new inputs of the same shape flow through the static buffers, and the device error word
    (bad species) is still raised."""
    from matten_b200.data.synthetic import synthetic_batch
    from matten_b200.graphs import CapturedForward

    dev = torch.device("cuda:0")
    _, prod = build_pair(HP_LMAX2, SPECIES8, torch.float32, dev, seed=5)
    b1 = _dev_batch(synthetic_batch(4, seed=0), dev, torch.float32)
    b2 = _dev_batch(synthetic_batch(4, seed=1), dev, torch.float32)
    with torch.no_grad():
        want1 = prod(dict(b1))["elastic_tensor_full"].clone()
        want2 = prod(dict(b2))["elastic_tensor_full"].clone()
        cf = CapturedForward(prod)
        got1 = cf(b1)["elastic_tensor_full"].clone()
        got2 = cf(b2)["elastic_tensor_full"].clone()
        got1b = cf(b1, slot=1)["elastic_tensor_full"].clone()
    assert torch.equal(got1, want1) and torch.equal(got2, want2) and torch.equal(got1b, want1)
    assert not torch.equal(want1, want2)
    bad = dict(b1)
    bad["atomic_numbers"] = b1["atomic_numbers"].clone()
    bad["atomic_numbers"][0] = 99
    with pytest.raises(RuntimeError, match="Invalid atomic numbers"):
        cf(bad)
    with torch.no_grad():
        assert torch.equal(cf(b1)["elastic_tensor_full"], want1)  # the flag is cleared by the next replay
