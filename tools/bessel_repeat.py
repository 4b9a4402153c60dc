"""Hunts the intermittent fp32 deviation of the Bessel radial basis.

Runs the kernel of ``test_edge_vectors_sh_radial`` many times on the test's fixed input and reports
  * whether any run differs bitwise from the first one (a race / stale buffer would show here),
  * the element-wise error of the GPU result against the fp64 evaluation of the same fp32 inputs,
  * the same error for the CPU fp32 oracle (torch's vectorised fp32 sin) -- the other suspect.
Usage: python tools/bessel_repeat.py [iterations]
"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matten_b200 import ops  # noqa: E402
from matten_b200.data.synthetic import synthetic_batch  # noqa: E402
from oracle import e3nn_restated as E  # noqa: E402
from oracle import matten_restated as M  # noqa: E402


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    dev = torch.device("cuda:0")
    b = synthetic_batch(3, dtype=torch.float32)
    ob = {k: v for k, v in b.items() if isinstance(v, torch.Tensor)}
    M.with_edge_vectors(ob)
    d = {k: v.to(dev) for k, v in b.items() if isinstance(v, torch.Tensor)}
    flag = ops.new_flag(dev)
    first = None
    n_diff = 0
    worst = 0.0
    for it in range(iters):
        # a fresh allocation pattern every iteration: garbage of different sizes in between
        junk = torch.full((1 + 7919 * (it % 13),), float("nan"), device=dev)
        _, ln = ops.edge_vectors(d["pos"], d["edge_index"], d["edge_cell_shift"], d["cell"], d["batch"], flag)
        emb = ops.edge_radial(ln, 0, 8, 0.0, 5.0, True)
        del junk
        ref64 = E.soft_one_hot_linspace_bessel(ln.cpu().double(), 0.0, 5.0, 8, True) * math.sqrt(8)
        err = float((emb.cpu().double() - ref64).abs().max() / ref64.abs().max())
        worst = max(worst, err)
        if first is None:
            first = emb.clone()
        elif not torch.equal(first, emb):
            n_diff += 1
    cpu32 = E.soft_one_hot_linspace_bessel(ob["edge_lengths"], 0.0, 5.0, 8, True) * math.sqrt(8)
    cpu64 = E.soft_one_hot_linspace_bessel(ob["edge_lengths"].double(), 0.0, 5.0, 8, True) * math.sqrt(8)
    cpu_err = float((cpu32.double() - cpu64).abs().max() / cpu64.abs().max())
    len_err = float((ln.cpu().double() - ob["edge_lengths"].double()).abs().max())
    print(json.dumps({"iters": iters, "runs_differing_bitwise_from_first": n_diff, "gpu_fp32_vs_fp64_max_normwise": worst,
                      "cpu_fp32_vs_fp64_max_normwise": cpu_err, "edge_len_gpu_vs_cpu_abs": len_err,
                      "cpu": torch.__config__.parallel_info().split("\n")[0], "threads": torch.get_num_threads()}))


if __name__ == "__main__":
    main()
