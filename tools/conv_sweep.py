"""BASELINE config 5: conv-layer sweep -- lmax 2/3/4, multiplicity 32..128, 1e5..1e7 edges, 32 in-edges per node,
senders uniform within blocks of 64 nodes, fp32.  Prints one JSON line per point: edges/s and the algorithmic-bytes
roofline fraction of SURVEY.md section 8(d).  Usage: python tools/conv_sweep.py [--quick] > profiles/conv_sweep.jsonl"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matten_b200 import o3  # noqa: E402
from matten_b200.graph import GraphCache  # noqa: E402
from matten_b200.nn.utils import UVUTensorProduct  # noqa: E402


def peaks():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def main():
    quick = "--quick" in sys.argv
    dev = torch.device("cuda:0")
    peak = peaks()
    g = torch.Generator().manual_seed(0)
    points = []
    for lmax in (2, 3, 4):
        for mul in (32, 64, 128):
            for E in (100_000, 1_000_000, 10_000_000):
                if E == 10_000_000 and (quick or mul > 32):
                    continue  # [E, (lmax+1)^2 * mul] gathers at 1e7 edges only for the narrow layers (time budget)
                points.append((lmax, mul, E))
    for lmax, mul, E in points:
        ir = o3.Irreps("+".join(f"{mul}x{l}{'e' if l % 2 == 0 else 'o'}" for l in range(lmax + 1)))
        sh = o3.Irreps.spherical_harmonics(lmax)
        tp = UVUTensorProduct(ir, sh, ir, mlp_input_size=8, mlp_hidden_size=32, mlp_num_hidden_layers=2,
                              mlp_activation="silu").to(dev)
        N = E // 32
        dst = torch.arange(N).repeat_interleave(32)
        src = (dst // 64) * 64 + torch.randint(0, 64, (E,), generator=g)
        src = src.clamp(max=N - 1)
        ei = torch.stack([src, dst]).to(dev)
        x = torch.randn(N, ir.dim, device=dev)
        v = torch.nn.functional.normalize(torch.randn(E, 3, device=dev), dim=1)
        from matten_b200 import ops
        y = ops.edge_sh(v, lmax, True)
        emb = torch.randn(E, 8, device=dev)
        gc = GraphCache({"edge_index": ei, "pos": torch.zeros(N, 3, device=dev)})
        with torch.no_grad():
            for _ in range(2):
                out = tp.fused(x, y, emb, gc, 32.0)
            torch.cuda.synchronize()
            reps = 5 if E <= 1_000_000 else 2
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(reps):
                out = tp.fused(x, y, emb, gc, 32.0)
            ev1.record()
            torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / reps
        pl = tp.plan
        W = pl.weight_numel
        nbytes = E * (8 + 4 * (pl.y_dim + 8 + pl.x_dim)) + N * 4 * pl.out_dim + 4 * (8 * 32 + 32 * 32 + 32 * W)
        print(json.dumps({"lmax": lmax, "mul": mul, "edges": E, "nodes": N, "paths": len(pl.paths), "weight_numel": W,
                          "x_dim": pl.x_dim, "D_mid": pl.out_dim, "kernel": "tcgen05" if pl.tc_num_tiles else "fma",
                          "ms": round(ms, 4), "edges_per_sec": round(E / ms * 1e3, 1),
                          "cg_macs_per_edge": pl.cg_macs_per_edge(), "algorithmic_MB": round(nbytes / 1e6, 1),
                          "achieved_GBps": round(nbytes / ms / 1e6, 1), "frac_of_hbm_peak": round(nbytes / ms / 1e6 / peak, 4),
                          "finite": bool(torch.isfinite(out).all())}), flush=True)
        del x, y, emb, out, gc, tp
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
