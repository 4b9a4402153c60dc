"""CPU check of the generated Clebsch-Gordan bundles (matten_b200/codegen/gen_bundles.py): the generated C++ is
compiled for the host with g++ (the CUDA qualifiers defined away) in double precision and every bundle's output
is compared with the dense einsum over the oracle's Wigner-3j tensors
(msg[m3] = sqrt(2 l3 + 1) w sum_{m1,m2} C x[m1] Y[m2]: SURVEY.md App. B, e3nn TensorProduct "uvu")."""
import math
import os
import shutil
import subprocess

import pytest
import torch

from matten_b200.codegen import gen_bundles
from oracle import e3nn_restated as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DRIVER = r"""
#include <cmath>
#include <cstdio>
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
using std::fma;
#include "cg_bundles.cuh"
using namespace mt;
template <int ID> void run(const double* x, const double* yrow, const double* w, unsigned mask) {
  using B = Bundle<ID>;
  double acc[B::NACC];
  for (int i = 0; i < B::NACC; ++i) acc[i] = 0;
  B::template edge<double>(x, yrow + B::Y_LO, w, acc, mask);
  std::printf("%d %d %d", ID, B::NACC, B::NP);
  for (int i = 0; i < B::NACC; ++i) std::printf(" %.17g", acc[i] * (double)B::scale(i));
  std::printf("\n");
}
int main() {
  double x[9], y[32], w[4];
  unsigned long long s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s % 2000001) / 1000000.0 - 1.0; };
  for (int i = 0; i < 9; ++i) x[i] = rnd();
  for (int i = 0; i < 32; ++i) y[i] = rnd();
  for (int i = 0; i < 4; ++i) w[i] = rnd();
#define X(ID) run<ID>(x, y, w, 0xffffffffu); run<ID>(x, y, w, 5u);
  MT_FOR_EACH_BUNDLE_L4(X)
  return 0;
}
"""


def _inputs():
    s = 88172645463325252
    out = []
    for _ in range(9 + 32 + 4):
        s ^= (s << 13) & 0xFFFFFFFFFFFFFFFF
        s ^= s >> 7
        s ^= (s << 17) & 0xFFFFFFFFFFFFFFFF
        out.append((s % 2000001) / 1000000.0 - 1.0)
    return out[:9], out[9:41], out[41:]


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_generated_bundles_match_the_dense_contraction(tmp_path):
    header_dir = os.path.join(ROOT, "matten_b200", "csrc", "generated")
    # the committed header must be what the generator produces now
    with open(os.path.join(header_dir, "cg_bundles.cuh")) as f:
        assert f.read() == gen_bundles.generate(), "run python -m matten_b200.codegen.gen_bundles"
    src = tmp_path / "drv.cpp"
    src.write_text(DRIVER)
    exe = tmp_path / "drv"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", header_dir, "-o", str(exe), str(src)], check=True)
    lines = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    x, y, w = _inputs()
    menu = gen_bundles.bundle_menu()
    assert len(lines) == 2 * len(menu)
    for i, b in enumerate(menu):
        for variant, mask in ((0, 0xFFFFFFFF), (1, 5)):
            tok = lines[2 * i + variant].split()
            assert int(tok[0]) == b.id and int(tok[1]) == b.n_acc and int(tok[2]) == len(b.paths)
            got = torch.tensor([float(t) for t in tok[3:]], dtype=torch.float64)
            want = torch.zeros(b.n_acc, dtype=torch.float64)
            for p, (l2, l3) in enumerate(b.paths):
                if not (mask >> p) & 1:
                    continue
                C = E.wigner_3j(b.l1, l2, l3).double() * math.sqrt(2 * l3 + 1)
                xv = torch.tensor(x[:2 * b.l1 + 1], dtype=torch.float64)
                yv = torch.tensor(y[gen_bundles.YPOS[l2]:gen_bundles.YPOS[l2] + 2 * l2 + 1], dtype=torch.float64)
                want[b.acc_off[p]:b.acc_off[p] + 2 * l3 + 1] = w[p] * torch.einsum("abc,a,b->c", C, xv, yv)
            assert torch.allclose(got, want, rtol=0, atol=2e-7), (b.id, b.paths, mask, (got - want).abs().max())


def test_menu_covers_every_path_up_to_lmax_4():
    seen = set()
    for b in gen_bundles.bundle_menu():
        assert b.n_acc <= max(gen_bundles.MAX_ACC, 9) and b.y_lo % 2 == 0 and b.y_cnt % 2 == 0
        for l2, l3 in b.paths:
            assert abs(b.l1 - l2) <= l3 <= b.l1 + l2
            assert (b.l1, l2, l3) not in seen
            seen.add((b.l1, l2, l3))
    want = {(a, b, c) for a in range(5) for b in range(5) for c in range(abs(a - b), min(4, a + b) + 1)}
    assert seen == want
    # covering a path list: masks select exactly the requested paths
    cov = gen_bundles.find_bundles(1, [(1, 0), (1, 2), (2, 2)])
    got = {(b.paths[i]) for b, m in cov for i in range(len(b.paths)) if (m >> i) & 1}
    assert got == {(1, 0), (1, 2), (2, 2)}
