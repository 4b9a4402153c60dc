"""predict() end to end on the GPU (BASELINE config 1 workload): checkpoint + config directory in the reference's
layout, example crystals, Cartesian [3,3,3,3] outputs against the oracle; failed entries and single-structure call."""
import json
import os

import numpy as np
import pytest
import torch
import yaml

from tests.helpers import GOLDEN, HP_LMAX4, build_pair, rel_err, to_oracle_batch

pytestmark = pytest.mark.gpu


def test_predict_matches_oracle_on_example_crystals(tmp_path):
    from matten_b200.data.neighbors import collate, make_graph
    from matten_b200.predict import predict, save_checkpoint

    dev = torch.device("cuda:0")
    with open(os.path.join(GOLDEN, "n100_structures.json")) as f:
        structs = json.load(f)["structures"][:10]
    species = sorted({z for s in structs for z in s["atomic_numbers"]})
    orac, prod = build_pair(HP_LMAX4, species, torch.float32, dev, seed=2)
    save_checkpoint(prod, tmp_path / "model_final.ckpt")
    with open(tmp_path / "config_final.yaml", "w") as f:
        yaml.safe_dump({"data": {"r_cut": 5.0, "tensor_target_name": "elastic_tensor_full",
                                 "tensor_target_formula": "ijkl=jikl=klij"}, "model": HP_LMAX4}, f)
    inputs = [{"lattice": s["lattice"], "atomic_numbers": s["atomic_numbers"], "cart_coords": s["cart_coords"]}
              for s in structs]
    # one structure that cannot be converted (a lone atom in a huge cell has no neighbours): None in the output
    inputs.insert(3, {"lattice": (np.eye(3) * 50).tolist(), "atomic_numbers": [species[0]], "cart_coords": [[0, 0, 0]]})
    with pytest.warns(UserWarning):
        preds = predict(inputs, model_identifier=str(tmp_path), batch_size=4, device=dev)
    assert len(preds) == 11 and preds[3] is None
    got = torch.as_tensor(np.stack([np.asarray(p) for i, p in enumerate(preds) if i != 3]))
    graphs = [make_graph(np.array(s["cart_coords"]), np.array(s["lattice"]), s["atomic_numbers"], 5.0, torch.float32)
              for s in structs]
    with torch.no_grad():
        want = orac.ct.to_cartesian(orac(to_oracle_batch(collate(graphs), torch.float32)))
    assert got.shape == (10, 3, 3, 3, 3)
    assert rel_err(got, want) < 1e-5
    # symmetric under ijkl = jikl = klij (reference tests/model/test_tfn_tensor.py)
    assert torch.allclose(got, got.permute(0, 2, 1, 3, 4), atol=1e-5 * float(got.abs().max()))
    assert torch.allclose(got, got.permute(0, 3, 4, 1, 2), atol=1e-5 * float(got.abs().max()))
    one = predict(inputs[0], model_identifier=str(tmp_path), device=dev)
    assert np.asarray(one).shape == (3, 3, 3, 3)
    assert np.allclose(np.asarray(one), got[0].numpy(), rtol=0, atol=1e-6 * float(got.abs().max()))
    bad = dict(inputs[0], atomic_numbers=[118] * len(inputs[0]["atomic_numbers"]))
    with pytest.raises(RuntimeError, match="not supported by the model"):
        predict(bad, model_identifier=str(tmp_path), device=dev)
