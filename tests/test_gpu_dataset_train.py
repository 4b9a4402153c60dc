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
