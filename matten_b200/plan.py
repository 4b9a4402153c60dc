"""
Init-time planners: turn irreps into the flat tables the CUDA kernels consume.

* :class:`UVUPlan`    -- the uvu instruction list of the reference
  (src/matten/nn/utils.py:205-237) regrouped into warp work items per (l1,l2,l3).
* :func:`linear_blocks` -- e3nn ``FullyConnectedTensorProduct(x, "Sx0e", out)`` /
  ``o3.Linear`` path lists as dense per-irrep blocks.
* :class:`GatePlan`   -- e3nn ``Gate`` (+ BatchNorm channel map) as per-output-element
  index tables.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch

from . import o3
from .codegen.gen_tables import LMAX, cg_nnz, cg_type_id
from .o3 import Irrep, Irreps


# --------------------------------------------------------------------------- #
@dataclass
class UVUPath:
    i_in1: int
    i_in2: int
    i_out: int  # index into the SORTED irreps_mid
    l1: int
    l2: int
    l3: int
    mul: int
    x_off: int
    y_off: int
    out_off: int
    w_off: int


class UVUPlan:
    """Paths + device tables of one uvu tensor product with per-edge weights."""

    def __init__(self, irreps_in1, irreps_in2, irreps_out):
        in1, in2, out = Irreps(irreps_in1), Irreps(irreps_in2), Irreps(irreps_out)
        self.irreps_in1, self.irreps_in2 = in1, in2
        mid, instr = [], []
        for i, (mul, ir1) in enumerate(in1):
            for j, (mul2, ir2) in enumerate(in2):
                if mul2 != 1:
                    raise NotImplementedError("uvu plan expects multiplicity-1 edge attributes")
                for ir_out in ir1 * ir2:
                    # the reference also tests `ir_out == Irreps("0e")`, which compares an Irrep
                    # with an Irreps and is always False (src/matten/nn/utils.py:210)
                    if ir_out in out:
                        instr.append((i, j, len(mid)))
                        mid.append((mul, ir_out))
        mid = Irreps(mid)
        if mid.dim == 0:
            raise ValueError(f"irreps_in1={in1} times irreps_in2={in2} produces no instructions in {out}")
        self.irreps_mid, perm, _ = mid.sort()
        s1, s2, so = in1.slices(), in2.slices(), self.irreps_mid.slices()
        self.paths: List[UVUPath] = []
        w_off = 0
        for i, j, k in instr:
            k = perm[k]
            mul, ir1 = in1[i]
            _, ir2 = in2[j]
            _, ir3 = self.irreps_mid[k]
            if max(ir1.l, ir2.l, ir3.l) > LMAX:
                raise NotImplementedError(f"l > {LMAX} is not generated (path {ir1} x {ir2} -> {ir3})")
            self.paths.append(UVUPath(i, j, k, ir1.l, ir2.l, ir3.l, mul, s1[i].start, s2[j].start,
                                      so[k].start, w_off))
            w_off += mul
        self.weight_numel = w_off
        self.x_dim, self.y_dim, self.out_dim = in1.dim, in2.dim, self.irreps_mid.dim
        self._build_items()
        self._build_tc()
        self._build_bwd()

    @property
    def irreps_out(self) -> Irreps:
        return self.irreps_mid.simplify()

    def _build_items(self):
        by_type: Dict[Tuple[int, int, int], List[Tuple[int, int, int, int]]] = {}
        for p in self.paths:
            cols = by_type.setdefault((p.l1, p.l2, p.l3), [])
            for u in range(p.mul):
                cols.append((p.w_off + u, p.x_off + u * (2 * p.l1 + 1), p.y_off, p.out_off + u * (2 * p.l3 + 1)))
        items = []  # (cost, type_id, cpw, [slots])
        for (l1, l2, l3), cols in by_type.items():
            tid = cg_type_id(l1, l2, l3)
            cost = cg_nnz(l1, l2, l3) + 2 * l3 + 1
            for c0 in range(0, len(cols), 32):
                chunk = cols[c0:c0 + 32]
                cpw = 1
                while cpw < len(chunk):
                    cpw *= 2
                slots = []
                for lane in range(32):
                    c = lane % cpw
                    slots.append(chunk[c] if c < len(chunk) else (-1, 0, 0, 0))
                items.append((cost, tid, cpw, slots))
        items.sort(key=lambda t: -t[0])  # heavy items first (dynamic scheduling)
        self.num_items = len(items)
        self.item_hdr = torch.tensor([[t[1], t[2]] for t in items], dtype=torch.int32)
        self.slot_tab = torch.tensor([t[3] for t in items], dtype=torch.int32).reshape(self.num_items, 32, 4)

    def _build_tc(self):
        """Tables of the tcgen05 path (csrc/conv_fwd_tc.cuh): see matten_b200/tcplan.py."""
        from .tcplan import TCPlan

        self.tc = TCPlan(self)

    @property
    def tc_num_tiles(self) -> int:
        """Total MMA tiles of the tcgen05 plan (0: the plan runs on the FMA-pipe kernel)."""
        return sum(p.num_tiles for p in self.tc.parts)

    def _build_bwd(self):
        """Tables of the backward kernel (csrc/conv_bwd.cuh).  The backward is organised by INPUT channel:
        a lane owns one channel (i, u) of ``x`` and walks every path that reads it, so the gradient of the
        gathered row is a register sum in a fixed path order (no atomics).
          bw_item_hdr [n,4]   : {l1, cpw, first path, path count}
          bw_lane_tab [n,32,2]: per lane {u (-1: idle), offset of x[u,:] in the x row}
          bw_path_tab [m,4]   : {cg_type_id, weight column of u = 0, sh offset, out offset of u = 0}"""
        by_in: Dict[int, List[UVUPath]] = {}
        for p in self.paths:
            by_in.setdefault(p.i_in1, []).append(p)
        items, path_tab = [], []
        s1 = self.irreps_in1.slices()
        for i, (mul, ir1) in enumerate(self.irreps_in1):
            ps = by_in.get(i, [])  # an input irrep without any path still gets its (zero) gradient written
            l1 = ir1.l
            if mul == 0 or l1 > LMAX:
                continue
            first = len(path_tab)
            for p in ps:
                path_tab.append([cg_type_id(p.l1, p.l2, p.l3), p.w_off, p.y_off, p.out_off])
            cost = sum(2 * cg_nnz(p.l1, p.l2, p.l3) + 2 * p.l3 + 2 for p in ps)
            for u0 in range(0, mul, 32):
                n = min(32, mul - u0)
                cpw = 1
                while cpw < n:
                    cpw *= 2
                lanes = []
                for lane in range(32):
                    c = lane % cpw
                    lanes.append([u0 + c, s1[i].start + (u0 + c) * (2 * l1 + 1)] if c < n else [-1, 0])
                items.append((cost, [l1, cpw, first, len(ps)], lanes))
        if not path_tab:
            path_tab.append([0, 0, 0, 0])
        items.sort(key=lambda t: -t[0])
        self.bw_num_items = len(items)
        self.bw_num_paths = len(path_tab)
        self.bw_item_hdr = torch.tensor([t[1] for t in items], dtype=torch.int32)
        self.bw_lane_tab = torch.tensor([t[2] for t in items], dtype=torch.int32).reshape(len(items), 32, 2)
        self.bw_path_tab = torch.tensor(path_tab, dtype=torch.int32).reshape(len(path_tab), 4)

    # per-edge algorithmic cost, for roofline reporting
    def cg_macs_per_edge(self) -> int:
        return sum(p.mul * (cg_nnz(p.l1, p.l2, p.l3) + 2 * p.l3 + 1) for p in self.paths)


# --------------------------------------------------------------------------- #
@dataclass
class LinBlock:
    in_off: int
    out_off: int
    mul_in: int
    mul_out: int
    dim: int
    w_off: int
    scale: float


def linear_blocks(irreps_in, irreps_out, num_species: int = 1) -> Tuple[List[LinBlock], int]:
    """Blocks + weight_numel of ``FullyConnectedTensorProduct(in, f"{S}x0e", out)``
    (num_species = S) or ``o3.Linear(in, out)`` (S = 1).  Weight of path (i_in, i_out) is
    flat ``[mul_in, S, mul_out]`` in instruction order (for i_in, for i_out); path scale
    ``1/sqrt(S * sum of mul_in feeding i_out)`` ('element' path normalisation)."""
    iin, iout = Irreps(irreps_in), Irreps(irreps_out)
    si, so = iin.slices(), iout.slices()
    instr = [(a, b) for a, (_, ir_a) in enumerate(iin) for b, (_, ir_b) in enumerate(iout) if ir_a == ir_b]
    blocks: List[LinBlock] = []
    w_off = 0
    for a, b in instr:
        mi, ir = iin[a]
        mo, _ = iout[b]
        fan = sum(iin[a2][0] for a2, b2 in instr if b2 == b) * num_species
        if mi > 0 and mo > 0:
            blocks.append(LinBlock(si[a].start, so[b].start, mi, mo, ir.dim, w_off, 1.0 / math.sqrt(fan)))
        w_off += mi * num_species * mo
    covered = {b for a, b in instr if iin[a][0] > 0}
    for b, (mo, ir) in enumerate(iout):
        if b not in covered and mo > 0:
            blocks.append(LinBlock(0, so[b].start, 0, mo, ir.dim, 0, 0.0))
    return blocks, w_off


# --------------------------------------------------------------------------- #
class GatePlan:
    """Irreps planning of reference ActivationLayer (src/matten/nn/utils.py:96-140) plus
    the element tables of e3nn Gate + BatchNorm for the fused kernel."""

    def __init__(self, tp_irreps_in1, tp_irreps_in2, tp_irreps_out, act_scalars: Dict[int, str],
                 act_gates: Dict[int, str]):
        out = Irreps(tp_irreps_out).sort().irreps.simplify()
        ok = lambda ir: tp_path_exists(tp_irreps_in1, tp_irreps_in2, ir)  # noqa: E731
        scalars = Irreps([(m, ir) for m, ir in out if ir.l == 0 and ok(ir)])
        gated = Irreps([(m, ir) for m, ir in out if ir.l > 0 and ok(ir)])
        if gated.dim > 0:
            if ok("0e"):
                g_ir = Irrep("0e")
            elif ok("0o"):
                g_ir = Irrep("0o")
            else:
                raise ValueError(f"tp_irreps_in1={tp_irreps_in1} times tp_irreps_in2={tp_irreps_in2} is unable "
                                 f"to produce gates needed for irreps_gated={gated}")
            gates = Irreps([(m, g_ir) for m, _ in gated]).simplify()
        else:
            gates = Irreps([])
        self.irreps_scalars, self.irreps_gates, self.irreps_gated = scalars, gates, gated
        self.irreps_in = (scalars + gates + gated).simplify()
        # output irreps: activations may flip the parity of odd scalars (e3nn.nn.Activation)
        out_scalars = []
        self._scalar_acts = []
        for m, ir in scalars:
            name = act_scalars[ir.p]
            f, aid = o3.ACT_FUNCS[name]
            p_out = ir.p
            if ir.p == -1:
                p_out = o3.act_parity(f)
                if p_out == 0:
                    raise ValueError("Activation: the parity is violated! The input scalar is odd but the "
                                     "activation is neither even nor odd.")
            out_scalars.append((m, (0, p_out)))
            self._scalar_acts.append((aid, o3.normalize2mom_const(f)))
        self._gate_acts = []
        gate_out = []
        for m, ir in gates:
            name = act_gates[ir.p]
            f, aid = o3.ACT_FUNCS[name]
            p_out = ir.p
            if ir.p == -1:
                p_out = o3.act_parity(f)
                if p_out == 0:
                    raise ValueError("Activation: the parity is violated!")
            gate_out.append((m, (0, p_out)))
            self._gate_acts.append((aid, o3.normalize2mom_const(f)))
        # gated (x) activated gates: elementwise product, parity multiplies
        gated_out = []
        gate_par = [p for m, (_, p) in gate_out for _ in range(m)]
        gi = 0
        for m, ir in gated:
            ps = set(gate_par[gi:gi + m])
            gi += m
            assert len(ps) == 1
            gated_out.append((m, (ir.l, ir.p * ps.pop())))
        self.irreps_out = Irreps(out_scalars) + Irreps(gated_out)
        self._build_tables()

    def _build_tables(self):
        ds, dg = self.irreps_scalars.dim, self.irreps_gates.dim
        src, gate, act, cst = [], [], [], []
        i = 0
        for (m, ir), (aid, c) in zip(self.irreps_scalars, self._scalar_acts):
            for _ in range(m):
                src.append(i)
                gate.append(-1)
                act.append(aid)
                cst.append(c)
                i += 1
        gate_act = []
        for (m, ir), (aid, c) in zip(self.irreps_gates, self._gate_acts):
            gate_act += [(aid, c)] * m
        pos = ds + dg
        gidx = 0
        for m, ir in self.irreps_gated:
            for u in range(m):
                aid, c = gate_act[gidx]
                for _ in range(ir.dim):
                    src.append(pos)
                    gate.append(ds + gidx)
                    act.append(aid)
                    cst.append(c)
                    pos += 1
                gidx += 1
        self.in_dim, self.out_dim = self.irreps_in.dim, self.irreps_out.dim
        assert pos == self.in_dim and len(src) == self.out_dim
        self.src_idx = torch.tensor(src, dtype=torch.int32)
        self.gate_idx = torch.tensor(gate, dtype=torch.int32)
        self.act_id = torch.tensor(act, dtype=torch.int32)
        self.act_cst = torch.tensor(cst, dtype=torch.float64)
        # inverse tables for the backward: input i feeds outputs inv_first[i] .. + inv_count[i] - 1
        inv_first, inv_count = [0] * self.in_dim, [0] * self.in_dim
        for j, (si, gi) in enumerate(zip(src, gate)):
            for i in ((si,) if gi < 0 else (si, gi)):
                if inv_count[i] == 0:
                    inv_first[i] = j
                assert j == inv_first[i] + inv_count[i], "outputs of one input must be contiguous"
                inv_count[i] += 1
        self.inv_first = torch.tensor(inv_first, dtype=torch.int32)
        self.inv_count = torch.tensor(inv_count, dtype=torch.int32)


def batchnorm_channel_map(irreps) -> Tuple[torch.Tensor, torch.Tensor, int, int]:
    """For e3nn BatchNorm(irreps): per element j of the feature row, the index of its
    channel in ``running_var``/``weight`` and (for 0e channels) in ``running_mean``/
    ``bias`` (else -1).  Returns (feat_idx, scalar_idx, num_features, num_scalar)."""
    irreps = Irreps(irreps)
    feat, scal = [], []
    f = s = 0
    for m, ir in irreps:
        for _ in range(m):
            for _ in range(ir.dim):
                feat.append(f)
                scal.append(s if ir.is_scalar() else -1)
            f += 1
            if ir.is_scalar():
                s += 1
    return (torch.tensor(feat, dtype=torch.int64), torch.tensor(scal, dtype=torch.int64), f, s)


def tp_path_exists(irreps_in1, irreps_in2, ir_out) -> bool:
    """reference src/matten/nn/utils.py:358-367"""
    a, b = Irreps(irreps_in1).simplify(), Irreps(irreps_in2).simplify()
    ir_out = Irrep(ir_out)
    return any(ir_out in ir1 * ir2 for _, ir1 in a for _, ir2 in b)
