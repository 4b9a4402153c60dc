"""Phase timing of the tcgen05 conv kernel (CTA 0, first 64 chunks) via in-kernel clock64 stamps."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matten_b200 import o3, ops  # noqa: E402
from matten_b200.graph import GraphCache  # noqa: E402
from matten_b200.nn.conv import PointConv  # noqa: E402

dev = torch.device("cuda:0")
layer = int(sys.argv[1]) if len(sys.argv) > 1 else 3
xin = ["16x0e", "32x0e+16x1o+4x2e", "32x0e+16x1o+16x1e+4x2o+4x2e", "32x0o+32x0e+16x1o+16x1e+4x2o+4x2e"][layer]
tgt = ["52x0e+16x1o+4x2e", "72x0e+16x1o+16x1e+4x2o+4x2e", "32x0o+72x0e+16x1o+16x1e+4x2o+4x2e",
       "32x0o+32x0e+16x1o+16x1e+4x2o+4x2e"][layer]
torch.manual_seed(0)
irreps_in = {"node_features": o3.Irreps(xin), "node_attrs": o3.Irreps("8x0e"),
             "edge_attrs": o3.Irreps.spherical_harmonics(2), "edge_embedding": o3.Irreps("8x0e")}
conv = PointConv(irreps_in, tgt, 2, 32, 28.0).to(dev)
N, deg = 32768, 28
E = N * deg
dst = torch.arange(N).repeat_interleave(deg)
src = (dst + torch.randint(1, 64, (E,))) % N
ei = torch.stack([src, dst]).to(dev)
x = torch.randn(N, conv.tp.plan.x_dim, device=dev)
sh = torch.randn(E, 9, device=dev)
emb = torch.randn(E, 8, device=dev)
data = {"edge_index": ei, "pos": torch.zeros(N, 3, device=dev)}
g = GraphCache(data)
dbg = torch.zeros(64 * 16 + 64 * 4, dtype=torch.int64, device=dev)
os.environ["MT_CONV_TC_DEBUG"] = str(dbg.data_ptr())
with torch.no_grad():
    for _ in range(3):
        out = conv.tp.fused(x, sh, emb, g, 28.0)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(10):
        out = conv.tp.fused(x, sh, emb, g, 28.0)
    ev1.record()
    torch.cuda.synchronize()
print("layer", layer, "ms per launch", ev0.elapsed_time(ev1) / 10, "items", conv.tp.plan.tc_num_sub)
u = dbg.cpu()[1024:].reshape(64, 4)
d = dbg.cpu()[:1024].reshape(64, 16)
names = ["wait_empty+meta", "bar", "issue_gather", "wait_bfree", "bar", "emb+wload", "cp_wait", "bar", "mlp", "pack",
         "bar", "mma_issue"]
print("chunk | " + " ".join(f"{n:>8s}" for n in ["waitE+m", "bar", "gather", "bfree", "bar", "emb+w", "cpwait", "bar", "mlp", "pack", "bar", "mma"])
      + " | total | cons: wait  work")
print('units (type, edges, cycles, cpw):', [tuple(int(v) for v in r) for r in u if r[2] > 0][:24])
for k in range(2, int(os.environ.get('TC_ROWS', '8'))):
    r = d[k]
    if r[0] == 0:
        break
    ph = [int(r[i + 1] - r[i]) for i in range(11)]
    tot = int(r[11] - r[0])
    cw, cwork = int(r[13] - r[12]), int(r[14] - r[13])
    print(f"{k:5d} | " + " ".join(f"{v:8d}" for v in ph) + f" | {tot:6d} | {cw:8d} {cwork:6d}")
