"""Phase timing of the tcgen05 conv kernel (CTA 0) via in-kernel clock64 accumulators (MT_CONV_TC_DEBUG)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matten_b200 import o3  # noqa: E402
from matten_b200.graph import GraphCache  # noqa: E402
from matten_b200.nn.conv import PointConv  # noqa: E402

dev = torch.device("cuda:0")
xins = ["16x0e", "32x0e+16x1o+4x2e", "32x0e+16x1o+16x1e+4x2o+4x2e", "32x0o+32x0e+16x1o+16x1e+4x2o+4x2e"]
tgts = ["52x0e+16x1o+4x2e", "72x0e+16x1o+16x1e+4x2o+4x2e", "32x0o+72x0e+16x1o+16x1e+4x2o+4x2e",
        "32x0o+32x0e+16x1o+16x1e+4x2o+4x2e"]
layers = [int(a) for a in sys.argv[1:]] or [0, 3]
for layer in layers:
    torch.manual_seed(0)
    irreps_in = {"node_features": o3.Irreps(xins[layer]), "node_attrs": o3.Irreps("8x0e"),
                 "edge_attrs": o3.Irreps.spherical_harmonics(2), "edge_embedding": o3.Irreps("8x0e")}
    conv = PointConv(irreps_in, tgts[layer], 2, 32, 28.0).to(dev)
    N, deg = 32768, 28
    E = N * deg
    dst = torch.arange(N).repeat_interleave(deg)
    src = (dst + torch.randint(1, 64, (E,))) % N
    ei = torch.stack([src, dst]).to(dev)
    x = torch.randn(N, conv.tp.plan.x_dim, device=dev)
    sh = torch.randn(E, 9, device=dev)
    emb = torch.randn(E, 8, device=dev)
    g = GraphCache({"edge_index": ei, "pos": torch.zeros(N, 3, device=dev)})
    dbg = torch.zeros(16384, dtype=torch.int64, device=dev)
    with torch.no_grad():
        for _ in range(3):
            conv.tp.fused(x, sh, emb, g, 28.0)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(10):
            conv.tp.fused(x, sh, emb, g, 28.0)
        ev1.record()
        torch.cuda.synchronize()
        os.environ["MT_CONV_TC_DEBUG"] = str(dbg.data_ptr())
        conv.tp.fused(x, sh, emb, g, 28.0)
        torch.cuda.synchronize()
        os.environ["MT_CONV_TC_DEBUG"] = ""
    d = dbg.cpu()[8192:]
    n = max(int(d[6]), 1)
    print(f"layer {layer}: {ev0.elapsed_time(ev1) / 10:.3f} ms per launch (untimed build), chunks of CTA 0: {n}")
    print(f"  warp0 cycles/chunk: wait_empty {int(d[0]) / n:.0f}, meta {int(d[1]) / n:.0f}, wait_bfree {int(d[2]) / n:.0f}, "
          f"pad+expect {int(d[3]) / n:.0f}, plane+sh copies {int(d[4]) / n:.0f}")
    print(f"  warp2/3 cycles/chunk: wait_go {int(d[10]) / n:.0f}/{int(d[12]) / n:.0f}, x copies {int(d[11]) / n:.0f}/{int(d[13]) / n:.0f}")
    print(f"  mma warp  cycles/chunk: wait_bready {int(d[8]) / n:.0f}, issue+commit {int(d[9]) / n:.0f}")
    cw = [(int(d[16 + 2 * w]) / n, int(d[17 + 2 * w]) / n) for w in range(4, 32)]
    for q in range(4):
        print(f"  consumers quarter {q} (wait_full, work) cycles/chunk: " +
              " ".join(f"({a:.0f},{b:.0f})" for w, (a, b) in enumerate(cw, start=4) if w % 4 == q))
    pl = conv.tp.plan
    hdr = pl.tc_sub_hdr.tolist()
    print("  sub-items (type, cpw, lane0, tile, quarter, model cost) -> measured cycles per unit:")
    for i, h in enumerate(hdr):
        cnt = int(d[256 + i])
        if cnt:
            print(f"    sub {i:2d} type {h[0]:2d} cpw {h[1]:2d} lane0 {h[2]:2d} tile {h[3]} q {h[4]}: {int(d[128 + i]) / cnt:8.0f}")
