"""A few points of the conv sweep (A/B experiments): python tools/sweep_points.py lmax,mul,E [...]"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matten_b200 import o3, ops
from matten_b200.graph import GraphCache
from matten_b200.nn.utils import UVUTensorProduct
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for spec in sys.argv[1:]:
    lmax, mul, E = [int(v) for v in spec.split(",")]
    ir = o3.Irreps("+".join(f"{mul}x{l}{'e' if l % 2 == 0 else 'o'}" for l in range(lmax + 1)))
    tp = UVUTensorProduct(ir, o3.Irreps.spherical_harmonics(lmax), ir, mlp_input_size=8, mlp_hidden_size=32,
                          mlp_num_hidden_layers=2, mlp_activation="silu").to(dev)
    N = E // 32
    dst = torch.arange(N).repeat_interleave(32)
    src = ((dst // 64) * 64 + torch.randint(0, 64, (E,), generator=g)).clamp(max=N - 1)
    ei = torch.stack([src, dst]).to(dev)
    x = torch.randn(N, ir.dim, device=dev)
    y = ops.edge_sh(torch.nn.functional.normalize(torch.randn(E, 3, device=dev), dim=1), lmax, True)
    emb = torch.randn(E, 8, device=dev)
    gc = GraphCache({"edge_index": ei, "pos": torch.zeros(N, 3, device=dev)})
    with torch.no_grad():
        for _ in range(3):
            tp.fused(x, y, emb, gc, 32.0)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(5):
            tp.fused(x, y, emb, gc, 32.0)
        ev1.record()
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 5
    print(spec, "tc" if tp.plan.tc_num_tiles else "fma", f"{ms:.3f} ms  {E / ms / 1e3:.1f}M e/s", flush=True)
