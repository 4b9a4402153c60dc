#!/usr/bin/env python
"""Training loop of the tensor-property models on the CUDA path (what scripts/train_materials_tensor.py and
scripts/train_atomic_tensor.py of the reference do through Lightning: MSE loss, Adam lr 0.01 / weight decay 1e-5,
optional data parallelism under torchrun with one NCCL all-reduce of the flat gradient per step).

    python scripts/train_tensor.py DATA.json [--atomic] [--epochs 2] [--batch-size 32] [--lmax 4]
    torchrun --nproc-per-node 8 scripts/train_tensor.py DATA.json ...
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matten_b200.dataset import TensorDataset  # noqa: E402
from matten_b200.model_factory import AtomicTensorModel, ScalarTensorModel  # noqa: E402
from matten_b200.predict import save_pretrained  # noqa: E402
from matten_b200.schedule import EarlyStopping, ReduceLROnPlateau  # noqa: E402
from matten_b200.train import Trainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("data")
    ap.add_argument("--atomic", action="store_true", help="per-atom rank-2 tensors (scripts/configs/atomic_tensor.yaml)")
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--batch-size", type=int, default=32)
    ap.add_argument("--lr", type=float, default=0.01)
    ap.add_argument("--weight-decay", type=float, default=1e-5)
    ap.add_argument("--out", default=None, help="write model_final.ckpt here")
    ap.add_argument("--val", default=None, help="validation file: enables ReduceLROnPlateau and early stopping on its MAE")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    if args.atomic:
        ds = TensorDataset(args.data, 5.0, "nmr_tensor", "irreps", "ij=ji", atom_selector="atom_selector", device=dev)
        hp = {"species_embedding_dim": 16, "irreps_edge_sh": "0e + 1o + 2e", "num_radial_basis": 8,
              "radial_basis_start": 0.0, "radial_basis_end": 5.0, "radial_basis_type": "bessel", "num_layers": 3,
              "invariant_layers": 2, "invariant_neurons": 32, "conv_layer_irreps": "32x0o+32x0e+16x1o+16x1e+4x2o+4x2e",
              "nonlinearity_type": "gate", "normalization": "batch", "resnet": True, "output_format": "irreps",
              "output_formula": "ij=ji", "conv_to_output_hidden_irreps_out": "16x0e + 2x2e", "reduce": "mean"}
        cls, key = AtomicTensorModel, "nmr_tensor"
    else:
        ds = TensorDataset(args.data, 5.0, "elastic_tensor_full", "irreps", "ijkl=jikl=klij", device=dev)
        hp = {"species_embedding_dim": 16, "irreps_edge_sh": "0e + 1o + 2e + 3o + 4e", "num_radial_basis": 8,
              "radial_basis_start": 0.0, "radial_basis_end": 5.0, "radial_basis_type": "bessel", "num_layers": 3,
              "invariant_layers": 2, "invariant_neurons": 32,
              "conv_layer_irreps": "32x0o+32x0e + 16x1o+16x1e + 4x2o+4x2e + 2x3o+2x3e + 2x4e",
              "nonlinearity_type": "gate", "normalization": "batch", "resnet": True,
              "conv_to_output_hidden_irreps_out": "16x0e + 2x2e + 4e", "output_format": "irreps",
              "output_formula": "ijkl=jikl=klij", "reduce": "mean"}
        cls, key = ScalarTensorModel, "elastic_tensor_full"
    hp["average_num_neighbors"] = ds.average_num_neighbors(dev)  # "auto" of the reference
    val = None
    if args.val:
        kw = dict(atom_selector="atom_selector") if args.atomic else {}
        val = TensorDataset(args.val, 5.0, key, "irreps", hp["output_formula"], device=dev, **kw)
    # species of every split (the reference takes them from the whole dataset's statistics)
    species = sorted(set(ds.species) | (set(val.species) if val is not None else set()))
    torch.manual_seed(3)
    model = cls(hp, {"allowed_species": species}, task_name=key).to(dev)
    trainer = Trainer(model, lr=args.lr, weight_decay=args.weight_decay, output_key=key)
    sched = ReduceLROnPlateau(trainer.opt, mode="min", factor=0.5, patience=50)  # config_final.yaml:17-23
    stopper = EarlyStopping(mode="min", patience=150)                            # materials_tensor.yaml:86-92
    for epoch in range(args.epochs):
        tot, n = 0.0, 0
        for batch, target, sel in ds.batches(args.batch_size, dev, shuffle=True, seed=epoch, rank=rank, world=world):
            loss = trainer.step(batch, target, atom_selector=sel)
            tot += float(loss)
            n += 1
        msg = f"epoch {epoch}: mean training loss {tot / max(n, 1):.6f} over {n} steps"
        if val is not None:
            m = trainer.evaluate(val.batches(args.batch_size, dev, rank=rank, world=world, even=False))
            sched.step(m["mae"])
            msg += f", val MAE {m['mae']:.6f}, lr {trainer.opt.lr:g}"
            if stopper.step(m["mae"]):
                if rank == 0:
                    print(msg + " -- early stop", flush=True)
                break
        if rank == 0:
            print(msg, flush=True)
    if args.out and rank == 0:
        # checkpoint + config_final.yaml: the directory is a valid `model_identifier` of matten_b200.predict.predict
        save_pretrained(model.eval(), args.out, r_cut=5.0, tensor_target_name=key, tensor_target_formula=hp["output_formula"])
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
