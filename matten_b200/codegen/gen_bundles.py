"""
Generates ``matten_b200/csrc/generated/cg_bundles.cuh``: the Clebsch-Gordan contractions of the
tcgen05 convolution (csrc/conv_fwd_tc.cuh), organised as *bundles*.

A bundle is a compile-time list of (l2, l3) paths of one input degree l1.  One warp unit runs a
bundle for 32 input channels (or 8 / 4 / 2 channels x edge phases): it loads the channel's x[u, :]
and the edge's spherical harmonics ONCE and feeds every path of the bundle, each path with its own
weight (same TMEM lane, another column range) and its own accumulators.  The reference reaches the
same arithmetic through e3nn's generated einsum per instruction (reference
src/matten/nn/utils.py:230-237 -> e3nn.o3.TensorProduct, instruction mode "uvu").

Instruction selection per path (every non-zero of the real Wigner-3j tensor is ONE fused
multiply-add, no separate multiply by the coefficient):
  * per output component c the most frequent |coefficient| s_c is pulled out of the edge loop
    (``SCALE``: applied once per node when the sum is stored);
  * the remaining ratios (1, 2, sqrt(3), ...) different from 1 use a pre-scaled operand, computed
    once per edge and shared by the non-zeros that need it;
  * signs are operand modifiers (free);
  * the per-edge weight enters through the cheapest of three forms:
        A: xw[a] = w x[a]          acc[c] += xw[a] y[b]
        B: t[c]  = sum x[a] y[b]   acc[c] += w t[c]
        C: yw[b] = w y[b]          acc[c] += x[a] yw[b]
All code is a template on the scalar type: ``f2`` (two edges per FFMA2) in the kernel, ``float`` /
``double`` for tests of the generated arithmetic.

Run:  python -m matten_b200.codegen.gen_bundles
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from functools import lru_cache
from typing import Dict, List, Tuple

from .. import o3

LMAX = 4
EPS = 1e-12
MAX_ACC = 12  # accumulators (sum of 2 l3 + 1) per bundle: x2 registers when two edges are packed

# position (in sh components) of the degree-l block inside the PADDED sh row the kernel keeps in shared
# memory: every block starts at an even position, so a pair-interleaved row can be read 16 bytes at a time
YPOS = [0, 2, 6, 12, 20]


def ypad(lmax: int) -> int:
    """Padded sh-row length (components) for harmonics up to lmax."""
    n = YPOS[lmax] + 2 * lmax + 1
    return n + (n & 1)


@dataclass
class Bundle:
    id: int
    l1: int
    paths: List[Tuple[int, int]]            # (l2, l3)
    acc_off: List[int] = field(default_factory=list)
    n_acc: int = 0
    y_lo: int = 0                           # first padded sh position the bundle reads (even)
    y_cnt: int = 0                          # number of padded positions (even)
    cost: List[int] = field(default_factory=list)    # floating-point instructions per edge (pair), per path
    scale: List[float] = field(default_factory=list)  # per accumulator

    @property
    def lmax(self) -> int:
        return max([self.l1] + [max(a, b) for a, b in self.paths])


def _split(l3s: List[int]) -> List[List[int]]:
    out, cur, tot = [], [], 0
    for l3 in l3s:
        d = 2 * l3 + 1
        if cur and tot + d > MAX_ACC:
            out.append(cur)
            cur, tot = [], 0
        cur.append(l3)
        tot += d
    if cur:
        out.append(cur)
    return out


@lru_cache(maxsize=None)
def bundle_menu() -> Tuple[Bundle, ...]:
    """The fixed menu.  l1 = 0: the paths (l2, l2) grouped {0,1,2}, {3}, {4} (they share x[u] and the whole
    sh row); l1 >= 1: per l2 the valid l3 in ascending order, cut so that a bundle keeps <= MAX_ACC
    accumulators."""
    menu: List[Bundle] = []

    def add(l1, paths):
        b = Bundle(len(menu), l1, list(paths))
        off = 0
        for _, l3 in paths:
            b.acc_off.append(off)
            off += 2 * l3 + 1
        b.n_acc = off
        lo = min(YPOS[l2] for l2, _ in paths)
        hi = max(YPOS[l2] + 2 * l2 + 1 for l2, _ in paths)
        b.y_lo = lo
        b.y_cnt = (hi - lo + 1) & ~1
        menu.append(b)

    add(0, [(0, 0), (1, 1), (2, 2)])
    add(0, [(3, 3)])
    add(0, [(4, 4)])
    for l1 in range(1, LMAX + 1):
        for l2 in range(LMAX + 1):
            l3s = list(range(abs(l1 - l2), min(LMAX, l1 + l2) + 1))
            for grp in _split(l3s):
                add(l1, [(l2, l3) for l3 in grp])
    for b in menu:
        for (l2, l3) in b.paths:
            code, cost, scale = _emit_path(b.l1, l2, l3, "x", f"(y + {YPOS[l2] - b.y_lo})", "w", "a")
            b.cost.append(cost)
            b.scale += scale
    return tuple(menu)


def find_bundles(l1: int, paths: List[Tuple[int, int]]) -> List[Tuple[Bundle, int]]:
    """Covers the (l2, l3) paths of one input degree with menu bundles: [(bundle, active-path bit mask)]."""
    want = set(paths)
    out = []
    for b in bundle_menu():
        if b.l1 != l1:
            continue
        mask = 0
        for i, p in enumerate(b.paths):
            if p in want:
                mask |= 1 << i
        if mask:
            out.append((b, mask))
            want -= set(b.paths)
    if want:
        raise NotImplementedError(f"no bundle for l1={l1} paths {sorted(want)}")
    return out


# --------------------------------------------------------------------------- #
def _lit(v: float) -> str:
    return f"T({v!r})"


def _emit_path(l1: int, l2: int, l3: int, X: str, Y: str, W: str, A: str):
    """C++ statements for one path: accumulates the UNSCALED sums into A[0 .. 2 l3], returns
    (lines, instruction count, [s_c])."""
    C = o3.wigner_3j(l1, l2, l3).double() * math.sqrt(2 * l3 + 1)
    d1, d2, d3 = 2 * l1 + 1, 2 * l2 + 1, 2 * l3 + 1
    by_c: Dict[int, List[Tuple[int, int, float]]] = {}
    for a in range(d1):
        for b in range(d2):
            for c in range(d3):
                v = C[a, b, c].item()
                if abs(v) > EPS:
                    by_c.setdefault(c, []).append((a, b, v))
    scale = []
    terms = []  # (c, a, b, sign, ratio)
    for c in range(d3):
        ent = by_c.get(c, [])
        if not ent:
            scale.append(0.0)
            continue
        mags: Dict[float, int] = {}
        for _, _, v in ent:
            key = round(abs(v), 10)
            mags[key] = mags.get(key, 0) + 1
        s_key = sorted(mags.items(), key=lambda kv: (-kv[1], kv[0]))[0][0]
        s = next(abs(v) for _, _, v in ent if round(abs(v), 10) == s_key)
        scale.append(s)
        for a, b, v in ent:
            r = abs(v) / s
            r = 1.0 if abs(r - 1.0) < 1e-9 else r
            terms.append((c, a, b, 1 if v > 0 else -1, r))
    used_a = sorted({a for _, a, _, _, _ in terms})
    used_b = sorted({b for _, _, b, _, _ in terms})
    sc_a = sorted({(a, round(r, 10)) for _, a, _, _, r in terms if r != 1.0})
    sc_b = sorted({(b, round(r, 10)) for _, _, b, _, r in terms if r != 1.0})
    nnz = len(terms)
    # scaled operands on the x side or on the y side, whichever needs fewer
    scale_on_x = len(sc_a) <= len(sc_b)
    n_sc = len(sc_a) if scale_on_x else len(sc_b)
    cost = {"A": nnz + len(used_a) + n_sc, "B": nnz + d3 + n_sc, "C": nnz + len(used_b) + n_sc}
    form = min(("A", "B", "C"), key=lambda f: (cost[f], f))
    L: List[str] = []
    rname = lambda r: f"{round(r * 1e6):d}"  # noqa: E731

    def xop(a, r):
        if form == "A":
            base = f"xw{a}"
        else:
            base = f"{X}[{a}]"
        if r != 1.0 and scale_on_x:
            return f"xs{a}_{rname(r)}"
        return base

    def yop(b, r):
        base = f"yw{b}" if form == "C" else f"{Y}[{b}]"
        if r != 1.0 and not scale_on_x:
            return f"ys{b}_{rname(r)}"
        return base

    if form == "A":
        for a in used_a:
            L.append(f"const T xw{a} = {W} * {X}[{a}];")
    if form == "C":
        for b in used_b:
            L.append(f"const T yw{b} = {W} * {Y}[{b}];")
    if scale_on_x:
        for a, r in sc_a:
            src = f"xw{a}" if form == "A" else f"{X}[{a}]"
            L.append(f"const T xs{a}_{rname(r)} = {_lit(r)} * {src};")
    else:
        for b, r in sc_b:
            src = f"yw{b}" if form == "C" else f"{Y}[{b}]"
            L.append(f"const T ys{b}_{rname(r)} = {_lit(r)} * {src};")
    if form == "B":
        for c in range(d3):
            ts = [t for t in terms if t[0] == c]
            if not ts:
                continue
            for i, (_, a, b, sg, r) in enumerate(ts):
                xo = xop(a, r)
                xo = xo if sg > 0 else f"(-{xo})"
                if i == 0:
                    L.append(f"T t{c} = {xo} * {yop(b, r)};")
                else:
                    L.append(f"t{c} = fma({xo}, {yop(b, r)}, t{c});")
            L.append(f"{A}[{c}] = fma({W}, t{c}, {A}[{c}]);")
    else:
        # round-robin over the output components: consecutive FMAs go to different accumulators (a lone warp
        # issues in order, back-to-back updates of one accumulator would each wait out the FMA latency)
        per_c = [[t for t in terms if t[0] == c] for c in range(d3)]
        order = []
        while any(per_c):
            for lst in per_c:
                if lst:
                    order.append(lst.pop(0))
        for (c, a, b, sg, r) in order:
            xo = xop(a, r)
            xo = xo if sg > 0 else f"(-{xo})"
            L.append(f"{A}[{c}] = fma({xo}, {yop(b, r)}, {A}[{c}]);")
    return L, cost[form], scale


def generate() -> str:
    menu = bundle_menu()
    L: List[str] = []
    L.append("// AUTO-GENERATED by matten_b200/codegen/gen_bundles.py -- do not edit by hand.")
    L.append("// Clebsch-Gordan bundles of the tcgen05 convolution: real Wigner-3j contractions (e3nn 0.5.x conventions,")
    L.append("// sqrt(2 l3 + 1) folded in), one fused multiply-add per non-zero, per-component scales applied per node.")
    L.append("#pragma once")
    L.append("namespace mt {")
    L.append(f"constexpr int kNumBundles = {len(menu)};")
    L.append(f"constexpr int kBundleMaxAcc = {max(b.n_acc for b in menu)};")
    L.append(f"constexpr int kBundleMaxPaths = {max(len(b.paths) for b in menu)};")
    L.append("// padded position of the degree-l block of the sh row kept in shared memory")
    L.append("__host__ __device__ constexpr int sh_pad_pos(int l) { return l == 0 ? 0 : l == 1 ? 2 : l == 2 ? 6 : l == 3 ? 12 : 20; }")
    L.append("__host__ __device__ constexpr int sh_pad_len(int lmax) { return lmax == 0 ? 2 : lmax == 1 ? 6 : lmax == 2 ? 12 : lmax == 3 ? 20 : 30; }")
    L.append("template <int ID> struct Bundle;")
    for b in menu:
        paths = ", ".join(f"({b.l1},{l2},{l3})" for l2, l3 in b.paths)
        L.append(f"// bundle {b.id}: l1 = {b.l1}, paths (l1,l2,l3) = {paths}; {sum(b.cost)} instructions per edge")
        L.append(f"template <> struct Bundle<{b.id}> {{")
        L.append(f"  static constexpr int L1 = {b.l1}, D1 = {2 * b.l1 + 1}, NP = {len(b.paths)}, NACC = {b.n_acc}, "
                 f"Y_LO = {b.y_lo}, Y_CNT = {b.y_cnt}, LMAXB = {b.lmax};")
        L.append("  // x: the channel's input components, y: sh components Y_LO .. Y_LO + Y_CNT - 1 of the padded row,")
        L.append("  // w: one weight per path, acc: NACC running (unscaled) sums, mask: active paths (warp uniform)")
        L.append("  template <typename T> static __device__ __forceinline__ void edge(const T* __restrict__ x, const T* __restrict__ y, "
                 "const T* __restrict__ w, T* __restrict__ acc, unsigned mask) {")
        for i, (l2, l3) in enumerate(b.paths):
            code, _, _ = _emit_path(b.l1, l2, l3, "x", f"(y + {YPOS[l2] - b.y_lo})", f"w[{i}]", f"(acc + {b.acc_off[i]})")
            L.append(f"    if (mask & {1 << i}u) {{  // ({b.l1},{l2},{l3})")
            for ln in code:
                L.append("      " + ln)
            L.append("    }")
        L.append("  }")
        L.append("  static __device__ __forceinline__ float scale(int i) {")
        L.append("    switch (i) {")
        for i, s in enumerate(b.scale):
            L.append(f"      case {i}: return {s!r}f;")
        L.append("      default: return 0.f;")
        L.append("    }")
        L.append("  }")
        L.append("  static __device__ __forceinline__ int path_d3(int p) { switch (p) { "
                 + " ".join(f"case {i}: return {2 * l3 + 1};" for i, (_, l3) in enumerate(b.paths)) + " default: return 0; } }")
        L.append("  static __device__ __forceinline__ int path_acc(int p) { switch (p) { "
                 + " ".join(f"case {i}: return {o};" for i, o in enumerate(b.acc_off)) + " default: return 0; } }")
        L.append("};")
    L.append("// X(bundle id): every bundle whose degrees are all <= 2 / <= 4")
    L.append("#define MT_FOR_EACH_BUNDLE_L2(X) \\")
    for b in menu:
        if b.lmax <= 2:
            L.append(f"  X({b.id}) \\")
    L.append("")
    L.append("#define MT_FOR_EACH_BUNDLE_L4(X) \\")
    for b in menu:
        L.append(f"  X({b.id}) \\")
    L.append("")
    L.append("#define MT_FOR_EACH_BUNDLE_GT2(X) \\")
    for b in menu:
        if b.lmax > 2:
            L.append(f"  X({b.id}) \\")
    L.append("")
    L.append("}  // namespace mt")
    return "\n".join(L) + "\n"


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    out = os.path.join(here, "..", "csrc", "generated", "cg_bundles.cuh")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    src = generate()
    with open(out, "w") as f:
        f.write(src)
    menu = bundle_menu()
    print(f"wrote {os.path.normpath(out)}: {len(menu)} bundles, {len(src.splitlines())} lines")


if __name__ == "__main__":
    main()
