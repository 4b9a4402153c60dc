"""
Planner of the tcgen05 convolution (csrc/conv_fwd_tc.cuh): turns the uvu path list of one tensor
product (reference src/matten/nn/utils.py:205-237) into the tables of the bundled kernel.

Vocabulary
  channel   one (input irrep i, multiplicity index u): the components x[u, :] of one sender row;
  group     <= 32 channels of one input degree l1 that one warp processes together:
              mode L: 32 channels, lane == channel, the warp walks every edge pair of a node;
              mode P: 8 / 4 / 2 channels (nch) x edge phases: thread t = 4 r + ph holds channel r % nch and
                      the edge pairs ph + 4 (r / nch) (mod 32 / nch) -- the tcgen05.ld.16x256b fragment;
  bundle    compile-time list of (l2, l3) paths of degree l1 that share the loads of x[u, :] and of the edge's
            spherical harmonics (matten_b200/codegen/gen_bundles.py);
  bundle instance (BI) = group x bundle x mask of the paths that exist in this tensor product.  It is the unit
            of work a consumer warp fetches (per receiver node).

The per-edge weights w[e, c] = MLP(emb_e)[c] are rows of the MMA accumulator in tensor memory: TMEM lane ==
row of the A operand (W_last^T), TMEM column == edge.  A warp reads only the 32 lanes 32 (warp % 4) .. of a
128-lane tile ("quarter"), so all weights of a BI live in ONE quarter, one 32-lane *slot* (mode L) or one
16-lane *half slot* per pair of paths (mode P) per path, in any of the part's tiles.

A tensor product whose weight rows do not fit 4 tiles (512 rows, the lmax-4 layers) is cut into *parts*:
disjoint sets of groups, each with its own tiles, run by its own CTAs over the same edges (every part
recomputes the small hidden layers of the radial MLP and gathers only its window of the x row).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import torch

from .codegen.gen_bundles import Bundle, find_bundles

MAX_PARTS = 4
MAX_TILES = 4          # 2 stages x tiles x chunk columns <= 512 TMEM columns
MAX_BI = 64            # bundle instances per part
REF_PAIRS = 14         # edge pairs per node of the cost model (28 neighbours)


@dataclass
class Chan:
    i_in: int
    u: int
    x_off: int                                   # offset of x[u, :] in the x row
    paths: Dict[Tuple[int, int], Tuple[int, int]]  # (l2, l3) -> (weight column, out offset of this channel)


@dataclass
class Group:
    l1: int
    mode: str            # "L" or "P"
    nch: int             # channels per thread row block: 32 (L) or 8 / 4 / 2 (P)
    chans: List[Chan]    # len <= nch (missing ones are dead channels)


@dataclass
class BI:
    group: Group
    bundle: Bundle
    mask: int
    cost: float = 0.0
    quarter: int = -1
    slots: List[int] = field(default_factory=list)   # per path slot: L: tile; P: (tile * 2 + half) per PAIR of paths

    def active(self, p: int) -> bool:
        return bool((self.mask >> p) & 1)

    @property
    def pairs(self) -> List[int]:
        """mode P: indices j of the path pairs (2 j, 2 j + 1) with at least one active path"""
        np_ = len(self.bundle.paths)
        return [j for j in range((np_ + 1) // 2) if self.active(2 * j) or (2 * j + 1 < np_ and self.active(2 * j + 1))]

    @property
    def need_half_slots(self) -> int:
        if self.group.mode == "L":
            return 2 * bin(self.mask).count("1")
        return len(self.pairs)


def _groups_of(chans: List[Chan], l1: int) -> List[Group]:
    out = []
    i = 0
    while len(chans) - i >= 32:
        out.append(Group(l1, "L", 32, chans[i:i + 32]))
        i += 32
    rest = chans[i:]
    while rest:
        n = 8 if len(rest) > 4 else (4 if len(rest) > 2 else 2)
        out.append(Group(l1, "P", n, rest[:n]))
        rest = rest[n:]
    return out


def _bi_cost(g: Group, b: Bundle, mask: int) -> float:
    """Issue slots of one (BI, node) unit for a node of REF_PAIRS edge pairs (relative)."""
    fp = sum(c for i, c in enumerate(b.cost) if (mask >> i) & 1)
    npaths = bin(mask).count("1")
    loads = 2 * (2 * b.l1 + 1) + b.y_cnt // 2
    per_iter = fp + loads + 4
    if g.mode == "L":
        return per_iter * REF_PAIRS + 0.75 * npaths * REF_PAIRS / 2 + 40
    phases = 32 // g.nch
    iters = -(-REF_PAIRS // phases)
    dup = 8 // g.nch
    return (per_iter + 6 + 3 * dup * ((npaths + 1) // 2)) * iters + 4 * b.n_acc * (2 + (dup - 1)) + 60


class TCPart:
    def __init__(self, bis: List[BI]):
        self.bis = bis
        self.cost = sum(b.cost for b in bis)
        self.num_tiles = 0


class TCPlan:
    """Tables of the tcgen05 path of one UVUPlan.  ``parts == []``: the plan does not qualify."""

    def __init__(self, uvu):
        self.parts: List[TCPart] = []
        self.y_lmax = max(ir.l for _, ir in uvu.irreps_in2)
        if self.y_lmax > 4 or any(mul != 1 for mul, _ in uvu.irreps_in2):
            return
        # sh blocks must be 0 .. y_lmax in order, one each (the kernel pads them by degree)
        if [ir.l for _, ir in uvu.irreps_in2] != list(range(self.y_lmax + 1)):
            return
        s1 = uvu.irreps_in1.slices()
        by_in: Dict[int, list] = {}
        for p in uvu.paths:
            by_in.setdefault(p.i_in1, []).append(p)
        chans_by_l1: Dict[int, List[Chan]] = {}
        for i, (mul, ir) in enumerate(uvu.irreps_in1):
            ps = by_in.get(i, [])
            if not ps or mul == 0:
                continue
            for u in range(mul):
                d = {}
                for p in ps:
                    if (p.l2, p.l3) in d:
                        return  # two paths of one channel with the same (l2, l3): not expressible as bundle masks
                    d[(p.l2, p.l3)] = (p.w_off + u, p.out_off + u * (2 * p.l3 + 1))
                chans_by_l1.setdefault(ir.l, []).append(Chan(i, u, s1[i].start + u * (2 * ir.l + 1), d))
        bis: List[BI] = []
        for l1 in sorted(chans_by_l1):
            for g in _groups_of(chans_by_l1[l1], l1):
                keys = sorted({k for c in g.chans for k in c.paths})
                for b, mask in find_bundles(l1, keys):
                    bi = BI(g, b, mask)
                    bi.cost = _bi_cost(g, b, mask)
                    bis.append(bi)
        if not bis:
            return
        # ---- parts: contiguous runs of BIs (they are ordered by l1, i.e. by x offset) with <= MAX_TILES tiles
        for nparts in range(1, MAX_PARTS + 1):
            parts = self._cut(bis, nparts)
            if parts is not None:
                self.parts = parts
                break
        if not self.parts:
            return
        self.x_dim, self.y_dim, self.out_dim = uvu.x_dim, uvu.y_dim, uvu.out_dim
        for part in self.parts:
            self._tables(part, uvu)

    # ------------------------------------------------------------------ packing
    @staticmethod
    def _pack(bis: List[BI], num_tiles: int) -> bool:
        """Longest-processing-time assignment of the BIs to the 4 quarters (= warp schedulers) under the slot
        capacity of ``num_tiles`` tiles; fills bi.quarter / bi.slots."""
        cap = [[[True, True] for _ in range(num_tiles)] for _ in range(4)]  # free half slots [q][tile][half]
        load = [0.0] * 4

        def take(q, bi: BI):
            got = []
            if bi.group.mode == "L":
                for p in range(len(bi.bundle.paths)):
                    if not bi.active(p):
                        got.append(-1)
                        continue
                    t = next((t for t in range(num_tiles) if cap[q][t][0] and cap[q][t][1]), None)
                    if t is None:
                        return None
                    cap[q][t][0] = cap[q][t][1] = False
                    got.append(t)
            else:
                for j in range((len(bi.bundle.paths) + 1) // 2):
                    if j not in bi.pairs:
                        got.append(-1)
                        continue
                    # prefer the second half of a tile whose first half is taken (keeps whole slots for mode L)
                    cand = [(t, h) for t in range(num_tiles) for h in (0, 1) if cap[q][t][h]]
                    if not cand:
                        return None
                    cand.sort(key=lambda th: (cap[q][th[0]][1 - th[1]], th[0], th[1]))
                    t, h = cand[0]
                    cap[q][t][h] = False
                    got.append(t * 2 + h)
            return got

        def free_halves(q):
            return sum(1 for t in range(num_tiles) for h in (0, 1) if cap[q][t][h])

        for bi in sorted(bis, key=lambda b: (-(b.need_half_slots if b.group.mode == "L" else 0), -b.cost)):
            order = sorted(range(4), key=lambda q: (load[q], -free_halves(q)))
            done = False
            for q in order:
                snap = [list(th) for th in cap[q]]
                got = take(q, bi)
                if got is not None:
                    bi.quarter, bi.slots = q, got
                    load[q] += bi.cost
                    done = True
                    break
                cap[q] = snap
            if not done:
                return False
        return True

    def _cut(self, bis: List[BI], nparts: int):
        """Best split of the BI list into ``nparts`` contiguous runs of whole groups (the BIs of a group share the
        gathered x window): the feasible split (every part packs into <= MAX_TILES tiles) with the smallest
        maximum part cost."""
        import itertools

        runs: List[List[BI]] = []
        for bi in bis:
            if runs and runs[-1][0].group is bi.group:
                runs[-1].append(bi)
            else:
                runs.append([bi])
        if nparts > len(runs):
            return None
        best = None
        for cuts in itertools.combinations(range(1, len(runs)), nparts - 1):
            bounds = (0,) + cuts + (len(runs),)
            chunks = [[bi for run in runs[bounds[i]:bounds[i + 1]] for bi in run] for i in range(nparts)]
            worst = max(sum(b.cost for b in c) for c in chunks)
            if best is not None and worst >= best[0]:
                continue
            if any(len(c) > MAX_BI or sum(b.need_half_slots for b in c) > 8 * MAX_TILES for c in chunks):
                continue
            best = (worst, chunks)
        if best is None:
            return None
        out = []
        for pb in best[1]:
            need = sum(b.need_half_slots for b in pb)
            for nt in range(max(1, -(-need // 8)), MAX_TILES + 1):
                if self._pack(pb, nt):
                    part = TCPart(pb)
                    part.num_tiles = nt
                    out.append(part)
                    break
            else:
                return None
        return out

    # ------------------------------------------------------------------ tables
    def _tables(self, part: TCPart, uvu):
        nt = part.num_tiles
        # quarters are interchangeable: number first the ones that reach into the last tile, so that the used A rows
        # end as early as possible (the kernel keeps only a_rows rows of the A operand in shared memory)
        def top_tile(q):
            t = [(s if bi.group.mode == "L" else s // 2) for bi in part.bis if bi.quarter == q for s in bi.slots if s >= 0]
            return max(t) if t else -1
        order = sorted(range(4), key=lambda q: (-top_tile(q), q))
        relabel = {q: i for i, q in enumerate(order)}
        for bi in part.bis:
            bi.quarter = relabel[bi.quarter]
        rows = [-1] * (nt * 128)
        hdr, lane_tab = [], []
        x_lo = min(c.x_off for bi in part.bis for c in bi.group.chans)
        x_lo -= x_lo % 4
        x_hi = max(c.x_off + 2 * bi.group.l1 + 1 for bi in part.bis for c in bi.group.chans)
        part.x_lo = x_lo
        part.x_cols = (x_hi - x_lo + 7) // 8 * 8   # TMA box: 4 gathered rows must be a multiple of 128 bytes
        part.lmax = max(bi.bundle.lmax for bi in part.bis)
        qlist = [[] for _ in range(4)]
        for k, bi in enumerate(part.bis):
            g, b = bi.group, bi.bundle
            q = bi.quarter
            slots3 = [0, 0, 0]
            lanes = [[0, -1, -1, -1] for _ in range(32)]
            for t in range(32):
                if g.mode == "L":
                    ci = t
                else:
                    ci = (t // 4) % g.nch
                if ci >= len(g.chans):
                    continue
                ch = g.chans[ci]
                lanes[t][0] = ch.x_off - x_lo
                stores = g.mode == "L" or (t % 4 == 0 and (t // 4) // g.nch == 0)
                for p, key in enumerate(b.paths):
                    if (bi.mask >> p) & 1 and key in ch.paths and stores:
                        lanes[t][1 + p] = ch.paths[key][1]
            for p, key in enumerate(b.paths):
                if not (bi.mask >> p) & 1:
                    continue
                if g.mode == "L":
                    tile = bi.slots[p]
                    slots3[p] = tile * 2
                    for t, ch in enumerate(g.chans):
                        if key in ch.paths:
                            rows[tile * 128 + 32 * q + t] = ch.paths[key][0]
                else:
                    s = bi.slots[p // 2]
                    tile, half = s // 2, s % 2
                    slots3[p] = s
                    for r8 in range(8):
                        ci = r8 % g.nch
                        if ci < len(g.chans) and key in g.chans[ci].paths:
                            rows[tile * 128 + 32 * q + 16 * half + 8 * (p % 2) + r8] = g.chans[ci].paths[key][0]
            hdr.append([b.id, 0 if g.mode == "L" else 1, g.nch, bi.mask, q, slots3[0], slots3[1], slots3[2]])
            lane_tab.append(lanes)
            qlist[q].append(k)
        for q in range(4):
            qlist[q].sort(key=lambda k: -part.bis[k].cost)
        last = max(i for i, r in enumerate(rows) if r >= 0)
        part.a_rows = (last // 32 + 1) * 32   # A rows kept in shared memory (the tail of the last tile is unused)
        part.row_wcol = torch.tensor(rows, dtype=torch.int32)
        part.bi_hdr = torch.tensor(hdr, dtype=torch.int32).reshape(len(hdr), 8)
        part.bi_lane = torch.tensor(lane_tab, dtype=torch.int32).reshape(len(hdr), 32, 4)
        ql = torch.zeros((4, MAX_BI), dtype=torch.int32)
        for q in range(4):
            if qlist[q]:
                ql[q, :len(qlist[q])] = torch.tensor(qlist[q], dtype=torch.int32)
        part.q_list = ql
        part.q_count = [len(qlist[q]) for q in range(4)]
        part.q_cost = [sum(part.bis[k].cost for k in qlist[q]) for q in range(4)]


# ------------------------------------------------------------------------- #
def emulate(plan: TCPlan, uvu, x: torch.Tensor, y: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """Host emulation of what the kernel computes from the tables for ONE edge: out [out_dim] from x [x_dim],
    y [y_dim], w [weight_numel] (fp64).  Used by the CPU tests to check the tables against the dense tensor
    product; mirrors the indexing of csrc/conv_fwd_tc.cuh (rows -> TMEM lanes, lane tables, masks, scales)."""
    import math

    from . import o3
    from .codegen.gen_bundles import bundle_menu

    out = torch.zeros(uvu.out_dim, dtype=torch.float64)
    written = torch.zeros(uvu.out_dim, dtype=torch.int32)
    menu = bundle_menu()
    ysl = uvu.irreps_in2.slices()
    for part in plan.parts:
        rows = part.row_wcol.tolist()
        tm = [w[r].item() if r >= 0 else 0.0 for r in rows]   # TMEM lane values of this edge, per tile row
        xw = x[part.x_lo:part.x_lo + part.x_cols] if part.x_lo + part.x_cols <= x.numel() else \
            torch.cat([x[part.x_lo:], torch.zeros(part.x_lo + part.x_cols - x.numel(), dtype=x.dtype)])
        for k in range(part.bi_hdr.shape[0]):
            bid, mode, nch, mask, q, s0, s1, s2 = part.bi_hdr[k].tolist()
            b = menu[bid]
            slots = [s0, s1, s2]
            for t in range(32):
                xoff, o0, o1, o2 = part.bi_lane[k, t].tolist()
                oo = [o0, o1, o2]
                for p, (l2, l3) in enumerate(b.paths):
                    if not (mask >> p) & 1 or oo[p] < 0:
                        continue
                    if mode == 0:
                        row = (slots[p] // 2) * 128 + 32 * q + t
                    else:
                        s = slots[p]
                        row = (s // 2) * 128 + 32 * q + 16 * (s % 2) + 8 * (p % 2) + (t // 4)
                    wv = tm[row]
                    C = o3.wigner_3j(b.l1, l2, l3).double() * math.sqrt(2 * l3 + 1)
                    xv = xw[xoff:xoff + 2 * b.l1 + 1].double()
                    yv = y[ysl[l2]].double()
                    out[oo[p]:oo[p] + 2 * l3 + 1] += wv * torch.einsum("abc,a,b->c", C, xv, yv)
                    written[oo[p]:oo[p] + 2 * l3 + 1] += 1
    assert int(written.min()) == 1 and int(written.max()) == 1, "every output element is produced exactly once"
    return out
