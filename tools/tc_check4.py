"""tcgen05 convolution vs the FMA-pipe kernel on the layers of the lmax-4 (pretrained 20230627) architecture."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matten_b200 import o3, ops  # noqa: E402
from matten_b200.graph import GraphCache  # noqa: E402
from matten_b200.nn.conv import PointConv  # noqa: E402

IR = "32x0o+32x0e+16x1o+16x1e+4x2o+4x2e+2x3o+2x3e+2x4e"
XINS = ["16x0e", "32x0e+16x1o+4x2e+2x3o+2x4e", "32x0o+32x0e+16x1o+16x1e+4x2o+4x2e+2x3o+2x3e+2x4e", IR]
TGTS = ["32x0e+16x1o+4x2e+2x3o+2x4e+24x0e", IR + "+28x0e", IR + "+28x0e", IR]


def main():
    dev = torch.device("cuda:0")
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    deg = int(sys.argv[2]) if len(sys.argv) > 2 else 28
    layers = [int(a) for a in sys.argv[3:]] or [0, 1, 2, 3]
    for layer in layers:
        torch.manual_seed(0)
        irreps_in = {"node_features": o3.Irreps(XINS[layer]), "node_attrs": o3.Irreps("8x0e"),
                     "edge_attrs": o3.Irreps.spherical_harmonics(4), "edge_embedding": o3.Irreps("8x0e")}
        conv = PointConv(irreps_in, TGTS[layer], 2, 32, 28.0).to(dev)
        pl = conv.tp.plan
        E = N * deg
        dst = torch.arange(N).repeat_interleave(deg)
        src = (dst + torch.randint(1, 64, (E,))) % N
        ei = torch.stack([src, dst]).to(dev)
        x = torch.randn(N, pl.x_dim, device=dev)
        sh = torch.randn(E, 25, device=dev)
        emb = torch.randn(E, 8, device=dev)
        g = GraphCache({"edge_index": ei, "pos": torch.zeros(N, 3, device=dev)})
        res = {"layer": layer, "x_dim": pl.x_dim, "W": pl.weight_numel, "D_mid": pl.out_dim, "N": N, "E": E,
               "parts": [(p.num_tiles, len(p.bis), p.x_lo, p.x_cols) for p in pl.tc.parts]}
        outs = {}
        with torch.no_grad():
            for impl in ("tc", "fma"):
                ops.conv_select_impl(impl)
                for _ in range(2):
                    out = conv.tp.fused(x, sh, emb, g, 28.0)
                torch.cuda.synchronize()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                for _ in range(3):
                    out = conv.tp.fused(x, sh, emb, g, 28.0)
                ev1.record()
                torch.cuda.synchronize()
                outs[impl] = out
                res[impl + "_ms"] = round(ev0.elapsed_time(ev1) / 3, 4)
            ops.conv_select_impl("auto")
        a, b = outs["tc"].double(), outs["fma"].double()
        res["normwise"] = float((a - b).abs().max() / b.abs().max())
        res["bad_elements"] = int(((a - b).abs() > 1e-4 * b.abs().max()).sum())
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
