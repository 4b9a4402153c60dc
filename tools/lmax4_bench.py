"""BASELINE config 3 flavour: the lmax-4 architecture of the pretrained 20230627 model (random weights) on synthetic
64-atom crystals -- inference forward and training step (fwd + bwd + Adam), batch 32 (the published batch size) and
512 crystals per GPU, fp32.  Prints one JSON line per point."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from matten_b200.model_factory import ScalarTensorModel  # noqa: E402
from matten_b200.train import Trainer  # noqa: E402

from tools.lmax4_hp import HP  # noqa: E402
dev = torch.device("cuda:0")
for ncry in (32, 512):
    torch.manual_seed(0)
    model = ScalarTensorModel(HP, {"allowed_species": bench.SPECIES}).to(dev)
    host = bench.make_batch(ncry, 0)
    keys = ["pos", "edge_index", "edge_cell_shift", "cell", "batch", "atomic_numbers", "num_neigh"]
    res = {k: host[k].to(dev) for k in keys}
    res["num_graphs"] = host["num_graphs"]
    E = host["edge_index"].shape[1]
    model.eval()
    with torch.no_grad():
        for _ in range(3):
            model(res, check=False)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(10):
            model(res, check=False)
        ev1.record()
        torch.cuda.synchronize()
    ms_fwd = ev0.elapsed_time(ev1) / 10
    tr = Trainer(model, lr=0.01, weight_decay=1e-5)
    target = torch.randn(ncry, 21, device=dev)
    for _ in range(3):
        tr.step(dict(res), target)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(5):
        loss = tr.step(dict(res), target)
    ev1.record()
    torch.cuda.synchronize()
    ms_tr = ev0.elapsed_time(ev1) / 5
    print(json.dumps({"model": "lmax-4 (pretrained 20230627 architecture), random weights, fp32", "crystals": ncry,
                      "atoms": int(host["pos"].shape[0]), "edges": int(E), "fwd_ms": round(ms_fwd, 3),
                      "fwd_crystals_per_s": round(ncry / ms_fwd * 1e3, 1), "train_step_ms": round(ms_tr, 3),
                      "train_crystals_per_s": round(ncry / ms_tr * 1e3, 1), "loss": round(float(loss), 5)}), flush=True)
    del tr, model
    torch.cuda.empty_cache()
