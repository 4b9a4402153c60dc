"""tcgen05 convolution vs the FMA-pipe kernel on the layers of the bench model: max deviation and time per call.
Usage: python tools/tc_check.py [N nodes] [degree] [layers ...]     (debugging / tuning aid)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matten_b200 import o3, ops  # noqa: E402
from matten_b200.graph import GraphCache  # noqa: E402
from matten_b200.nn.conv import PointConv  # noqa: E402

XINS = ["16x0e", "32x0e+16x1o+4x2e", "32x0e+16x1o+16x1e+4x2o+4x2e", "32x0o+32x0e+16x1o+16x1e+4x2o+4x2e"]
TGTS = ["52x0e+16x1o+4x2e", "72x0e+16x1o+16x1e+4x2o+4x2e", "32x0o+72x0e+16x1o+16x1e+4x2o+4x2e",
        "32x0o+32x0e+16x1o+16x1e+4x2o+4x2e"]


def main():
    dev = torch.device("cuda:0")
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    deg = int(sys.argv[2]) if len(sys.argv) > 2 else 28
    layers = [int(a) for a in sys.argv[3:]] or [0, 1, 2, 3]
    for layer in layers:
        torch.manual_seed(0)
        irreps_in = {"node_features": o3.Irreps(XINS[layer]), "node_attrs": o3.Irreps("8x0e"),
                     "edge_attrs": o3.Irreps.spherical_harmonics(2), "edge_embedding": o3.Irreps("8x0e")}
        conv = PointConv(irreps_in, TGTS[layer], 2, 32, 28.0).to(dev)
        E = N * deg
        dst = torch.arange(N).repeat_interleave(deg)
        src = (dst + torch.randint(1, 64, (E,))) % N
        ei = torch.stack([src, dst]).to(dev)
        x = torch.randn(N, conv.tp.plan.x_dim, device=dev)
        sh = torch.randn(E, 9, device=dev)
        emb = torch.randn(E, 8, device=dev)
        g = GraphCache({"edge_index": ei, "pos": torch.zeros(N, 3, device=dev)})
        res = {"layer": layer, "N": N, "E": E, "parts": [(p.num_tiles, len(p.bis)) for p in conv.tp.plan.tc.parts]}
        outs = {}
        with torch.no_grad():
            for impl in ("tc", "fma"):
                ops.conv_select_impl(impl)
                for _ in range(2):
                    out = conv.tp.fused(x, sh, emb, g, 28.0)
                torch.cuda.synchronize()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                for _ in range(5):
                    out = conv.tp.fused(x, sh, emb, g, 28.0)
                ev1.record()
                torch.cuda.synchronize()
                outs[impl] = out
                res[impl + "_ms"] = round(ev0.elapsed_time(ev1) / 5, 4)
            ops.conv_select_impl("auto")
        if os.environ.get("TC_TIMING"):
            from matten_b200 import _lib
            import ctypes
            dbg = torch.zeros(32 * 8, dtype=torch.int64, device=dev)
            _lib.load().mt_conv_set_debug_buffer(ctypes.c_void_p(dbg.data_ptr()))
            ops.conv_select_impl("tc")
            with torch.no_grad():
                conv.tp.fused(x, sh, emb, g, 28.0)
            torch.cuda.synchronize()
            _lib.load().mt_conv_set_debug_buffer(None)
            ops.conv_select_impl("auto")
            t = dbg.cpu().reshape(32, 8)
            for w in range(20):
                role = ("prod[layout,wait_empty,wait_bfree,copies]" if w == 0 else "mma[wait_bready,issue]" if w == 1
                        else "gather[wait_go,issue]" if w in (2, 3) else "cons[wait_full,units,..,n_units]")
                print(f"   warp {w:2d} {role}: " + " ".join(f"{int(v) / 1e3:.0f}k" if i < 7 else str(int(v)) for i, v in enumerate(t[w].tolist())), flush=True)
        a, b = outs["tc"].double(), outs["fma"].double()
        res["max_abs_diff"] = float((a - b).abs().max())
        res["max_ref"] = float(b.abs().max())
        res["normwise"] = res["max_abs_diff"] / res["max_ref"]
        bad = ((a - b).abs() > 1e-4 * res["max_ref"]).nonzero()
        res["bad_elements"] = int(bad.shape[0])
        if bad.shape[0]:
            res["first_bad"] = [bad[0].tolist(), float(a[tuple(bad[0])]), float(b[tuple(bad[0])])]
            res["bad_cols"] = sorted(set(bad[:, 1].tolist()))[:40]
            res["bad_rows"] = sorted(set(bad[:, 0].tolist()))[:20]
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
