"""Host-side planning of the product (matten_b200/o3.py, plan.py) against the independent oracle
restatement, and the bookkeeping numbers of SURVEY.md Appendix A."""
import math

import numpy as np
import pytest
import torch

from matten_b200 import o3
from matten_b200.codegen.gen_tables import cg_nnz, cg_types
from matten_b200.plan import GatePlan, UVUPlan, linear_blocks
from oracle import e3nn_restated as E
from oracle import matten_restated as M
from tests.helpers import HP_LMAX2, HP_LMAX4, HP_REFTEST, SPECIES8


def test_irreps_algebra():
    ir = o3.Irreps("32x0o+32x0e + 16x1o+16x1e + 4x2o+4x2e")
    assert ir.dim == 200 and ir.num_irreps == 104 and ir.lmax == 2
    assert str(o3.Irreps("1e + 0e + 1e").sort().irreps) == "1x0e+1x1e+1x1e"
    assert o3.Irreps("2o + 1e + 0e + 1e").sort().p == (3, 1, 0, 2)
    assert str(o3.Irreps("0e+0o").sort().irreps) == "1x0o+1x0e"
    assert str((o3.Irreps("3x0e") + o3.Irreps("2x0e+1o")).simplify()) == "5x0e+1x1o"
    assert [str(i) for i in o3.Irrep("1o") * o3.Irrep("2e")] == ["1o", "2o", "3o"]
    assert o3.Irrep("2e") in o3.Irreps("4x2e") and o3.Irrep("2o") not in o3.Irreps("4x2e")
    assert str(o3.Irreps.spherical_harmonics(2)) == "1x0e+1x1o+1x2e"
    assert (o3.Irrep(0, 1) == o3.Irreps("0e")) is False  # the always-False clause of utils.py:210


def test_wigner_matches_oracle():
    for l1, l2, l3 in cg_types():
        assert (o3.wigner_3j(l1, l2, l3) - E.wigner_3j(l1, l2, l3)).abs().max() < 1e-14


def test_cg_nnz_table():
    # SURVEY.md App. A
    expect = {(0, 0, 0): 1, (0, 1, 1): 3, (1, 1, 1): 6, (1, 1, 2): 11, (1, 2, 2): 16, (1, 2, 3): 21, (2, 2, 2): 25,
              (2, 2, 4): 37, (3, 3, 3): 42, (3, 3, 4): 67, (4, 4, 4): 97, (2, 3, 4): 50, (4, 4, 2): 61}
    for k, v in expect.items():
        assert cg_nnz(*k) == v, k
    assert len(cg_types()) == 65


@pytest.mark.parametrize("formula", ["ijkl=jikl=klij", "ij=ji", "ij", "ij=-ji", "ijk=jik"])
def test_cartesian_tensor_matches_oracle(formula):
    ct = o3.CartesianTensor(formula)
    ir, Q = E.reduced_tensor_products(formula)
    assert o3.parse_irreps_list(ct) == ir
    assert np.abs(ct.change_of_basis(torch.float64).numpy() - Q.reshape(Q.shape[0], -1)).max() < 1e-12


def test_normalize2mom_matches_oracle():
    import torch.nn.functional as F

    assert o3.normalize2mom_const(F.silu) == E.normalize2mom(F.silu).cst
    assert o3.normalize2mom_const(torch.tanh) == E.normalize2mom(torch.tanh).cst
    assert o3.normalize2mom_const(torch.sigmoid) == E.normalize2mom(torch.sigmoid).cst


@pytest.mark.parametrize("hp,expect", [
    (HP_LMAX2, [(3, 48, 144, 288), (15, 216, 696, 2096), (27, 336, 1104, 3616), (30, 432, 1392, 4192)]),
    (HP_LMAX4, [(5, 80, 400, 800), (59, 452, 2324, 9854), (99, 714, 3658, 16632), (103, 842, 4170, 17656)]),
    (HP_REFTEST, [(5, 160, 800, None), (65, 608, 3328, None), (125, 1056, 5856, None), (130, 1216, 6656, None)]),
])
def test_uvu_plan_matches_appendix_a_and_oracle(hp, expect):
    from matten_b200.model_factory import ScalarTensorModel

    species = SPECIES8 if hp is not HP_REFTEST else [8, 52]
    prod = ScalarTensorModel(hp, {"allowed_species": species})
    orac = M.ScalarTensorModel(hp, {"allowed_species": species})
    got = []
    for (n1, m1), (n2, m2) in zip(prod.backbone.named_children(), orac.backbone.named_children()):
        assert n1 == n2
        p1, p2 = getattr(m1, "conv", m1), getattr(m2, "conv", m2)
        if not hasattr(p1, "tp"):
            continue
        plan = p1.tp.plan
        got.append((len(plan.paths), plan.weight_numel, plan.out_dim, plan.cg_macs_per_edge()))
        # same instructions, same sorted mid irreps, same offsets as the oracle's e3nn-style TP
        tp = p2.tp.tp
        assert o3.parse_irreps_list(plan.irreps_mid) == tp.irreps_out
        assert [(p.i_in1, p.i_in2, p.i_out) for p in plan.paths] == [(i.i_in1, i.i_in2, i.i_out)
                                                                      for i in tp.instructions]
        assert all(abs(i.path_weight - math.sqrt(2 * tp.irreps_out[i.i_out][1] + 1)) < 1e-12
                   for i in tp.instructions)
        assert o3.parse_irreps_list(m1.irreps_out["node_features"]) == \
            [tuple(t) for t in m2.irreps_out["node_features"]]
        # every weight column appears exactly once in the slot table
        cols = plan.slot_tab[:, :, 0].reshape(-1)
        cols = cols[cols >= 0]
        per_item_cols = [set(plan.slot_tab[i, :, 0][plan.slot_tab[i, :, 0] >= 0].tolist())
                         for i in range(plan.num_items)]
        assert sorted(set().union(*per_item_cols)) == list(range(plan.weight_numel))
        assert sum(len(s) for s in per_item_cols) == plan.weight_numel
    for g, e in zip(got, expect):
        assert g[:3] == e[:3]
        if e[3] is not None:
            assert g[3] == e[3]
    # state_dict keys/shapes are interchangeable
    sd_p, sd_o = prod.state_dict(), orac.state_dict()
    for k, v in sd_p.items():
        assert k in sd_o and tuple(sd_o[k].shape) == tuple(v.shape), k


def test_linear_blocks_scales():
    blocks, numel = linear_blocks("54x0o+56x0e+100x1o", "32x0o+32x0e+16x1o+4x2e", 8)
    assert numel == (54 * 32 + 56 * 32 + 100 * 16) * 8
    real = [b for b in blocks if b.mul_in > 0]
    assert [round(b.scale, 12) for b in real] == [round(1 / math.sqrt(8 * m), 12) for m in (54, 56, 100)]
    zero = [b for b in blocks if b.mul_in == 0]
    assert len(zero) == 1 and zero[0].mul_out == 4 and zero[0].dim == 5
    # o3.Linear with two inputs feeding one output: fan-in is summed
    blocks, numel = linear_blocks("4x0e+6x0e", "3x0e", 1)
    assert numel == 30 and all(abs(b.scale - 1 / math.sqrt(10)) < 1e-12 for b in blocks)


def test_gate_plan_irreps():
    sh = "0e+1o+2e+3o+4e"
    g = GatePlan("16x0e", sh, HP_LMAX4["conv_layer_irreps"], {1: "silu", -1: "tanh"}, {1: "sigmoid", -1: "tanh"})
    assert str(g.irreps_in) == "56x0e+16x1o+4x2e+2x3o+2x4e"
    assert g.irreps_in.dim == 156 and g.irreps_out.dim == 132
    g = GatePlan("32x0e+16x1o+4x2e+2x3o+2x4e", sh, HP_LMAX4["conv_layer_irreps"], {1: "silu", -1: "tanh"},
                 {1: "sigmoid", -1: "tanh"})
    assert g.irreps_in.dim == 260 and g.irreps_out.dim == 214
    g = GatePlan(g.irreps_out, sh, HP_LMAX4["conv_layer_irreps"], {1: "silu", -1: "tanh"},
                 {1: "sigmoid", -1: "tanh"})
    assert str(g.irreps_in) == "32x0o+78x0e+16x1o+16x1e+4x2o+4x2e+2x3o+2x3e+2x4e"
    assert g.irreps_in.dim == 292 and g.irreps_out.dim == 246


@pytest.mark.parametrize("x_ir,sh_lmax,out_ir", [
    ("16x0e", 2, "52x0e+16x1o+4x2e"),
    ("32x0o+32x0e+16x1o+16x1e+4x2o+4x2e", 2, "32x0o+32x0e+16x1o+16x1e+4x2o+4x2e"),
    ("32x0o+32x0e+16x1o+16x1e+4x2o+4x2e+2x3o+2x3e+2x4e", 4, "32x0o+32x0e+16x1o+16x1e+4x2o+4x2e+2x3o+2x3e+2x4e"),
    ("20x0e+12x1o+5x2e+3x1e", 2, "20x0e+12x1o+5x2e+3x1e"),
])
def test_backward_tables_cover_every_channel_and_weight_column_once(x_ir, sh_lmax, out_ir):
    """Tables of mt_conv_bwd (plan._build_bwd): every element of the x row belongs to exactly one (item, lane) --
    the gradient of the gathered row is written, never accumulated -- and every weight column / output slot of the
    forward plan is reached through exactly one (item, lane, path)."""
    from matten_b200 import o3
    from matten_b200.codegen.gen_tables import cg_type_id
    from matten_b200.plan import UVUPlan

    pl = UVUPlan(o3.Irreps(x_ir), o3.Irreps.spherical_harmonics(sh_lmax), o3.Irreps(out_ir))
    hdr, lanes, paths = pl.bw_item_hdr.tolist(), pl.bw_lane_tab.tolist(), pl.bw_path_tab.tolist()
    tid2l = {cg_type_id(p.l1, p.l2, p.l3): (p.l1, p.l2, p.l3) for p in pl.paths}
    x_seen = [0] * pl.x_dim
    w_seen = [0] * pl.weight_numel
    out_seen = [0] * pl.out_dim
    for (l1, cpw, first, count), lt in zip(hdr, lanes):
        assert 32 % cpw == 0
        for lane in range(cpw):  # the other lanes repeat these channels on further edge phases
            u, xoff = lt[lane]
            if u < 0:
                continue
            for m in range(2 * l1 + 1):
                x_seen[xoff + m] += 1
            for tid, wcol0, yoff, ooff0 in paths[first:first + count]:
                a, b, c = tid2l[tid]
                assert a == l1 and yoff == b * b
                w_seen[wcol0 + u] += 1
                for m in range(2 * c + 1):
                    out_seen[ooff0 + u * (2 * c + 1) + m] += 1
        for lane in range(cpw, 32):
            assert lt[lane] == lt[lane % cpw]
    assert x_seen == [1] * pl.x_dim
    assert w_seen == [1] * pl.weight_numel
    assert out_seen == [1] * pl.out_dim


@pytest.mark.parametrize("x_ir,out_ir,lmax", [
    ("16x0e", "52x0e+16x1o+4x2e", 2),
    ("32x0e+16x1o+4x2e", "72x0e+16x1o+16x1e+4x2o+4x2e", 2),
    ("32x0o+32x0e+16x1o+16x1e+4x2o+4x2e", "32x0o+32x0e+16x1o+16x1e+4x2o+4x2e", 2),
    ("20x0e+12x1o+4x2e+4x1e", "20x0e+12x1o+4x2e+4x1e", 2),
    ("32x0e+32x1o+32x2e", "32x0e+32x1o+32x2e", 2),
    ("32x0o+32x0e+16x1o+16x1e+4x2o+4x2e+2x3o+2x3e+4x4e", "32x0o+32x0e+16x1o+16x1e+4x2o+4x2e+2x3o+2x3e+4x4e", 4),
    ("8x0e+8x1o+8x2e+8x3o+8x4e", "8x0e+8x1o+8x2e+8x3o+8x4e", 4),
])
def test_tcgen05_tables_reproduce_the_tensor_product(x_ir, out_ir, lmax):
    """Tables of the tensor-core path (matten_b200/tcplan.py): every weight column sits on a row of the MMA A operand
    (= one TMEM lane of one tile; duplicated only over the edge-phase row blocks of mode P), all rows of a bundle
    instance lie in its own 32-lane quarter, no slot is used twice -- and a host emulation of the kernel's indexing
    (rows -> lanes, lane tables, path masks, per-part x windows) reproduces the dense uvu tensor product of the
    oracle for one edge, every output element written exactly once."""
    from matten_b200 import o3
    from matten_b200.codegen.gen_bundles import bundle_menu
    from matten_b200.plan import UVUPlan
    from matten_b200.tcplan import MAX_TILES, emulate
    from oracle import e3nn_restated as E

    pl = UVUPlan(o3.Irreps(x_ir), o3.Irreps.spherical_harmonics(lmax), o3.Irreps(out_ir))
    tc = pl.tc
    assert tc.parts, "plan must qualify for the tensor-core path"
    menu = bundle_menu()
    seen_cols = set()
    for part in tc.parts:
        assert 1 <= part.num_tiles <= MAX_TILES and part.x_lo % 4 == 0 and part.x_cols % 8 == 0
        rows = part.row_wcol.tolist()
        assert len(rows) == part.num_tiles * 128
        owner = {}
        for k, (bid, mode, nch, mask, q, s0, s1, s2) in enumerate(part.bi_hdr.tolist()):
            b = menu[bid]
            assert mode in (0, 1) and nch in ((32,) if mode == 0 else (8, 4, 2)) and 0 <= q < 4 and mask
            for p_i in range(len(b.paths)):
                if not (mask >> p_i) & 1:
                    continue
                s = (s0, s1, s2)[p_i]
                tile, half = s // 2, s % 2
                assert 0 <= tile < part.num_tiles
                lanes = range(32) if mode == 0 else range(16 * half + 8 * (p_i % 2), 16 * half + 8 * (p_i % 2) + 8)
                for ln in lanes:
                    key = (tile, q, ln)
                    assert key not in owner, "two paths share a TMEM lane"
                    owner[key] = (k, p_i)
                if mode == 1:  # duplicate row blocks hold the same weight column
                    base = tile * 128 + 32 * q + 16 * half + 8 * (p_i % 2)
                    for r8 in range(8):
                        assert rows[base + r8] == rows[base + r8 % nch]
        for r, wc in enumerate(rows):
            if wc >= 0:
                assert (r // 128, (r % 128) // 32, r % 32) in owner
                seen_cols.add(wc)
        assert sum(part.q_count) == part.bi_hdr.shape[0]
        ql = part.q_list.tolist()
        listed = sorted(s for q in range(4) for s in ql[q][:part.q_count[q]])
        assert listed == list(range(part.bi_hdr.shape[0]))
    assert seen_cols == set(range(pl.weight_numel))
    # one edge through the tables vs the oracle's dense tensor product
    torch.manual_seed(1)
    x = torch.randn(pl.x_dim, dtype=torch.float64)
    y = torch.randn(pl.y_dim, dtype=torch.float64)
    w = torch.randn(pl.weight_numel, dtype=torch.float64)
    got = emulate(tc, pl, x, y, w)
    instr = [(p.i_in1, p.i_in2, p.i_out, "uvu", True) for p in pl.paths]
    tp = E.TensorProduct(str(pl.irreps_in1), str(pl.irreps_in2), str(pl.irreps_mid), instr,
                         shared_weights=False, internal_weights=False).double()
    want = tp(x[None], y[None], w[None])[0]
    assert torch.allclose(got, want, rtol=0, atol=1e-12), float((got - want).abs().max())
