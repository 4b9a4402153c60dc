"""
Generates the small fixtures under tests/golden/ (run in the build container, where
/root/reference exists; the GPU box only sees the committed outputs).

  elastic_tensor_one_structure.json : lattice / coordinates / atomic numbers of the reference's own
      test fixture tests/test_files/elastic_tensor_one.json (the 8-atom TeO3 cell of its equivariance
      test, tests/model/test_tfn_tensor.py:106) plus its DFT elastic tensor.
  n100_structures.json : the 100 crystals of datasets/example_crystal_elasticity_tensor_n100.json
      (lattice, coords, Z) -- the workload of BASELINE config 1.
  oracle_lmax2_seed0.pt : inputs + fp64 outputs of the ORACLE (not of e3nn: e3nn cannot be installed
      here, see oracle/e3nn_restated.py) for a 4-crystal synthetic batch; a regression pin.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

_Z = {s: i + 1 for i, s in enumerate(
    "H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr Rb Sr Y Zr "
    "Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir "
    "Pt Au Hg Tl Pb Bi Po At Rn Fr Ra Ac Th Pa U Np Pu".split())}


def _struct(s):
    lat = np.array(s["lattice"]["matrix"])
    frac = np.array([site["abc"] for site in s["sites"]])
    z = [_Z[site["species"][0]["element"]] for site in s["sites"]]
    return {"lattice": lat.tolist(), "cart_coords": (frac @ lat).tolist(), "atomic_numbers": z}


def main():
    d = json.load(open(os.path.join(REF, "tests/test_files/elastic_tensor_one.json")))
    out = _struct(d["structure"]["0"])
    out["elastic_tensor_full"] = d["elastic_tensor_full"]["0"]
    json.dump(out, open(os.path.join(HERE, "elastic_tensor_one_structure.json"), "w"))

    d = json.load(open(os.path.join(REF, "datasets/example_crystal_elasticity_tensor_n100.json")))
    keys = sorted(d["structure"].keys(), key=int)
    structs = [_struct(d["structure"][k]) for k in keys]
    json.dump({"structures": structs}, open(os.path.join(HERE, "n100_structures.json"), "w"))
    print("n100:", len(structs), "crystals,", sum(len(s["atomic_numbers"]) for s in structs), "atoms,",
          len({z for s in structs for z in s["atomic_numbers"]}), "elements")

    # small slices of the two example datasets IN THE REFERENCE'S OWN FILE FORMAT (pandas-style column JSON with
    # pymatgen Structure dicts): inputs of the dataset-reader / training tests
    keep = keys[:6]
    json.dump({c: {k: d[c][k] for k in keep} for c in ("structure", "elastic_tensor_full")},
              open(os.path.join(HERE, "elasticity_n6_reference_format.json"), "w"))
    n = json.load(open(os.path.join(REF, "datasets/si_nmr_data.json")))
    nk = [k for k in sorted(n["structure"].keys(), key=int) if len(n["species"][k]) <= 30][:4]
    json.dump({c: {k: n[c][k] for k in nk} for c in ("structure", "nmr_tensor", "atom_selector")},
              open(os.path.join(HERE, "si_nmr_n4_reference_format.json"), "w"))

    from matten_b200.data.synthetic import synthetic_batch
    from oracle import matten_restated as M
    from tests.helpers import HP_LMAX2, SPECIES8, randomize_bn

    torch.manual_seed(0)
    torch.set_default_dtype(torch.float64)
    model = M.ScalarTensorModel(HP_LMAX2, {"allowed_species": SPECIES8})
    randomize_bn(model, 0)
    # weights rounded to fp32-representable values so the fixture can store them as fp32
    model.load_state_dict({k: (v.float().double() if v.is_floating_point() else v)
                           for k, v in model.state_dict().items()})
    model.eval()
    batch = synthetic_batch(2, dtype=torch.float64)
    with torch.no_grad():
        d2 = {k: v for k, v in batch.items() if isinstance(v, torch.Tensor)}
        feats = model.backbone(dict(d2))
        out = model(d2)
    sd = {k: (v.float() if v.is_floating_point() else v) for k, v in model.state_dict().items()
          if "output_mask" not in k}
    torch.save({"state_dict": sd, "batch": {k: v for k, v in batch.items()},
                "node_features": feats["node_features"], "output": out},
               os.path.join(HERE, "oracle_lmax2_seed0.pt"))
    print("oracle golden:", tuple(out.shape), float(out.abs().max()))


if __name__ == "__main__":
    main()
