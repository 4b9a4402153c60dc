"""Host logic of predict(): structure parsing, species check, tolerant checkpoint loading (Lightning checkpoints
reference classes of packages that are not installed here)."""
import pickle

import numpy as np
import pytest
import torch

from matten_b200 import predict as P


def test_structure_arrays_from_dict_variants():
    lat = np.array([[3.348898, 0.0, 1.933487], [1.116299, 3.157372, 1.933487], [0.0, 0.0, 3.866975]])
    frac = [[0.25, 0.25, 0.25], [0, 0, 0]]
    a = P._structure_arrays({"lattice": lat, "species": ["Si", "Si"], "coords": frac})
    assert a["Z"] == [14, 14]
    assert np.allclose(a["cart"], np.asarray(frac) @ lat)
    b = P._structure_arrays({"lattice": {"matrix": lat.tolist()}, "atomic_numbers": [14, 14],
                             "cart_coords": (np.asarray(frac) @ lat).tolist()})
    assert np.allclose(a["cart"], b["cart"]) and b["Z"] == [14, 14]
    # pymatgen Structure.as_dict() layout
    c = P._structure_arrays({"lattice": {"matrix": lat.tolist()},
                             "sites": [{"species": [{"element": "Si", "occu": 1}], "xyz": x.tolist()}
                                       for x in np.asarray(frac) @ lat]})
    assert np.allclose(a["cart"], c["cart"]) and c["Z"] == [14, 14]


def test_check_species_message():
    class M:
        hparams = {"dataset_hparams": {"allowed_species": [8, 14]}}

    P.check_species(M(), [{"Z": [14, 8]}])
    with pytest.raises(RuntimeError, match=r"structure 1.*Fe \(26\).*not supported"):
        P.check_species(M(), [{"Z": [14]}, {"Z": [26, 8]}])


class _Gone:  # stands for e.g. a torchmetrics object pickled into a Lightning checkpoint
    def __init__(self):
        self.value = 3


def test_load_checkpoint_with_missing_classes(tmp_path):
    import sys
    import types

    mod = types.ModuleType("not_installed_pkg")
    cls = type("Metric", (), {"__module__": "not_installed_pkg"})
    mod.Metric = cls
    sys.modules["not_installed_pkg"] = mod
    obj = cls()
    obj.value = 3
    ck = {"state_dict": {"backbone.w": torch.arange(4.0)}, "hyper_parameters": {"dataset_hparams": {"allowed_species": [1]}},
          "callbacks": {"m": obj}}
    path = tmp_path / "model_final.ckpt"
    torch.save(ck, path)
    del sys.modules["not_installed_pkg"]  # the package is "not installed" at load time
    out = P.load_checkpoint(path)
    assert torch.equal(out["state_dict"]["backbone.w"], torch.arange(4.0))
    assert out["hyper_parameters"]["dataset_hparams"]["allowed_species"] == [1]
    # a bare state_dict file is wrapped
    torch.save({"w": torch.ones(2)}, tmp_path / "sd.pt")
    assert "state_dict" in P.load_checkpoint(tmp_path / "sd.pt")


def test_predict_requires_cuda_when_absent():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.predict({"lattice": np.eye(3) * 4, "species": ["Si"], "coords": [[0, 0, 0]]})
