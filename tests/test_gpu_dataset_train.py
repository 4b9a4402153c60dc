"""Dataset reader (the reference's column-JSON format) + training loop on the GPU: targets are the reference's
``from_cartesian`` projection (checked against the oracle's CartesianTensor), ``average_num_neighbors="auto"``,
a few Adam steps reduce the loss for the pooled elasticity model and for the per-atom NMR model with its atom selector."""
import json
import os

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN, HP_LMAX2, HP_LMAX4, rel_err

pytestmark = pytest.mark.gpu


def test_elasticity_dataset_targets_and_training():
    from matten_b200.dataset import TensorDataset
    from matten_b200.model_factory import ScalarTensorModel
    from matten_b200.train import Trainer
    from oracle import e3nn_restated as E

    dev = torch.device("cuda:0")
    path = os.path.join(GOLDEN, "elasticity_n6_reference_format.json")
    ds = TensorDataset(path, 5.0, "elastic_tensor_full", "irreps", "ijkl=jikl=klij", device=dev, dtype=torch.float64)
    assert len(ds) == 6 and ds.targets[0].shape == (1, 21)
    raw = json.load(open(path))
    ct = E.CartesianTensor("ijkl=jikl=klij")
    for i, k in enumerate(raw["elastic_tensor_full"]):
        want = ct.from_cartesian(torch.as_tensor(np.asarray(raw["elastic_tensor_full"][k], dtype=np.float64)))
        assert rel_err(ds.targets[i].reshape(-1), want.reshape(-1)) < 1e-12
    avg = ds.average_num_neighbors(dev)
    assert 10.0 < avg < 80.0
    hp = dict(HP_LMAX4, average_num_neighbors=avg)
    torch.manual_seed(0)
    model = ScalarTensorModel(hp, {"allowed_species": ds.species}).to(dev)
    scale = float(torch.cat(ds.targets).abs().max())
    tr = Trainer(model, lr=0.01, weight_decay=1e-5)
    losses = []
    for epoch in range(8):
        for batch, target, sel in ds.batches(6, dev, torch.float32):
            assert sel is None and target.shape == (6, 21)
            losses.append(float(tr.step(batch, (target / scale).float())))
    assert all(np.isfinite(losses)) and losses[-1] < 0.7 * losses[0], losses


def test_atomic_nmr_dataset_with_selector_trains():
    from matten_b200.dataset import TensorDataset
    from matten_b200.model_factory import AtomicTensorModel
    from matten_b200.train import Trainer

    dev = torch.device("cuda:0")
    ds = TensorDataset(os.path.join(GOLDEN, "si_nmr_n4_reference_format.json"), 5.0, "nmr_tensor", "irreps", "ij=ji",
                       atom_selector="atom_selector", device=dev)
    assert len(ds) == 4 and {8, 14} <= set(ds.species)
    nsel = sum(int(s.sum()) for s in ds.selectors)
    assert sum(t.shape[0] for t in ds.targets) == nsel and ds.targets[0].shape[1] == 6
    hp = dict(HP_LMAX2, average_num_neighbors=ds.average_num_neighbors(dev))
    hp.pop("conv_to_output_hidden_irreps_out", None)
    torch.manual_seed(0)
    model = AtomicTensorModel(hp, {"allowed_species": ds.species}).to(dev)
    tr = Trainer(model, lr=0.01, weight_decay=1e-5, output_key="nmr_tensor")
    scale = float(torch.cat(ds.targets).abs().max())
    losses = []
    for epoch in range(8):
        for batch, target, sel in ds.batches(4, dev):
            assert int(sel.sum()) == target.shape[0]
            losses.append(float(tr.step(batch, target / scale, atom_selector=sel)))
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses


def test_edgeless_crystals_are_skipped_and_evaluation_sees_every_batch(tmp_path):
    """Reference behaviour (dataset/structure_scalar_tensor.py:357-362): a crystal without any edge inside r_cut is
    skipped and recorded; evaluation batches are not truncated to a multiple of the world size."""
    from matten_b200.dataset import TensorDataset

    dev = torch.device("cuda:0")
    raw = json.load(open(os.path.join(GOLDEN, "elasticity_n6_reference_format.json")))
    keys = list(raw["structure"].keys())
    lonely = json.loads(json.dumps(raw["structure"][keys[0]]))
    # one atom in a 30 A cubic box: no neighbour within 5 A, not even a periodic image
    lonely["lattice"]["matrix"] = [[30.0, 0, 0], [0, 30.0, 0], [0, 0, 30.0]]
    for k in ("a", "b", "c"):
        if k in lonely["lattice"]:
            lonely["lattice"][k] = 30.0
    lonely["sites"] = lonely["sites"][:1]
    lonely["sites"][0]["abc"] = [0.0, 0.0, 0.0]
    lonely["sites"][0]["xyz"] = [0.0, 0.0, 0.0]
    raw["structure"]["lonely"] = lonely
    raw["elastic_tensor_full"]["lonely"] = raw["elastic_tensor_full"][keys[0]]
    path = tmp_path / "with_lonely.json"
    path.write_text(json.dumps(raw))
    with pytest.warns(UserWarning, match="Skipped 1 structures"):
        ds = TensorDataset(str(path), 5.0, "elastic_tensor_full", "irreps", "ijkl=jikl=klij", device=dev)
    assert len(ds) == 6 and ds.failed_entries == ["lonely"] and len(ds.targets) == 6
    # 6 crystals in batches of 2 = 3 chunks over 2 ranks: training drops the odd chunk, evaluation keeps it
    n_train = [len(list(ds.batches(2, dev, rank=r, world=2))) for r in (0, 1)]
    n_eval = [len(list(ds.batches(2, dev, rank=r, world=2, even=False))) for r in (0, 1)]
    assert n_train == [1, 1] and n_eval == [2, 1]
