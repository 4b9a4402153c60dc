import sys, os, torch
sys.path.insert(0, os.getcwd())
import bench
from matten_b200.model_factory import ScalarTensorModel
from matten_b200.train import Trainer
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = ScalarTensorModel(bench.HP, {"allowed_species": bench.SPECIES}).to(dev)
host = bench.make_batch(512, 0)
keys = ["pos", "edge_index", "edge_cell_shift", "cell", "batch", "atomic_numbers", "num_neigh"]
res = {k: host[k].to(dev) for k in keys}; res["num_graphs"] = host["num_graphs"]
tr = Trainer(m)
tgt = torch.randn(512, 6, device=dev)
for _ in range(2): tr.step(dict(res), tgt)
torch.cuda.synchronize()
