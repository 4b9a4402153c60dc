"""
``torch.library`` custom ops over the C ABI: namespace ``matten_b200``.

The module layer (``matten_b200.nn``) reaches the kernels through ``matten_b200.functional`` /
``matten_b200.autograd`` (``torch.autograd.Function``); the same entry points are registered here as dispatcher-level
custom ops -- tensors and plain numbers only in the schemas, shape inference on fake / meta tensors, autograd formulas
that call the backward kernels -- so that code which keeps the reference's own modules can call e.g.
``torch.ops.matten_b200.conv_fwd`` directly (option B) and so that the ops are visible to
``torch.compile`` / export as opaque calls.

Plans (immutable host-side tables + their device copies) are Python objects; ops refer to them by an integer id
obtained from ``register_plan`` (``plan_id`` arguments).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor

from . import ops

_PLANS: Dict[int, object] = {}


def register_plan(handle) -> int:
    """Make a ``ConvPlanHandle`` / ``LinPlanHandle`` addressable from op schemas; returns its id."""
    pid = id(handle)
    _PLANS[pid] = handle
    return pid


def _plan(pid: int):
    try:
        return _PLANS[pid]
    except KeyError:
        raise RuntimeError(f"matten_b200: unknown plan id {pid}; call torch_ops.register_plan(handle) first") from None


# ------------------------------------------------------------------------------------------ geometry / bookkeeping
@torch.library.custom_op("matten_b200::edge_sh", mutates_args=())
def edge_sh(edge_vec: Tensor, lmax: int, normalize: bool) -> Tensor:
    return ops.edge_sh(edge_vec, lmax, normalize)


@edge_sh.register_fake
def _(edge_vec, lmax, normalize):
    return edge_vec.new_empty((edge_vec.shape[0], (lmax + 1) ** 2))


@torch.library.custom_op("matten_b200::edge_radial", mutates_args=())
def edge_radial(edge_len: Tensor, mode: int, num_basis: int, start: float, end: float, cutoff: bool,
                poly_p: float) -> Tensor:
    return ops.edge_radial(edge_len, mode, num_basis, start, end, cutoff, poly_p, None)


@edge_radial.register_fake
def _(edge_len, mode, num_basis, start, end, cutoff, poly_p):
    return edge_len.new_empty((edge_len.shape[0], num_basis))


@torch.library.custom_op("matten_b200::csr_by_key", mutates_args=())
def csr_by_key(keys: Tensor, num_keys: int) -> Tuple[Tensor, Tensor]:
    rowptr, perm = ops.csr_by_key(keys, num_keys, True, None)
    return rowptr, perm


@csr_by_key.register_fake
def _(keys, num_keys):
    return (keys.new_empty((num_keys + 1,), dtype=torch.int32), keys.new_empty((keys.shape[0],), dtype=torch.int32))


@torch.library.custom_op("matten_b200::segment_reduce", mutates_args=())
def segment_reduce(x: Tensor, ptr: Tensor, reduce: str) -> Tensor:
    return ops.segment_reduce(x, ptr, reduce)


@segment_reduce.register_fake
def _(x, ptr, reduce):
    return x.new_empty((ptr.shape[0] - 1, x.shape[1]))


def _segment_reduce_bwd(ctx, g):
    return ops.segment_reduce_bwd(g.contiguous(), ctx.ptr, ctx.n, ctx.reduce), None, None


def _segment_reduce_setup(ctx, inputs, output):
    x, ptr, reduce = inputs
    ctx.ptr, ctx.n, ctx.reduce = ptr, x.shape[0], reduce


segment_reduce.register_autograd(_segment_reduce_bwd, setup_context=_segment_reduce_setup)


# ------------------------------------------------------------------------------------------ fused convolution
@torch.library.custom_op("matten_b200::conv_fwd", mutates_args=())
def conv_fwd(x: Tensor, sh: Tensor, emb: Tensor, mlp_weights: List[Tensor], rowptr: Tensor, perm: Tensor,
             src_sorted: Tensor, plan_id: int, avg_num_neighbors: float, num_neigh: Optional[Tensor]) -> Tensor:
    """radial MLP -> uvu tensor product -> receiver sum -> / sqrt(#neighbours)  (``mt_conv_fwd``);
    ``avg_num_neighbors <= 0`` selects the per-node ``num_neigh`` normalisation."""
    return ops.conv_fwd(_plan(plan_id), x, sh, emb, list(mlp_weights), rowptr, perm, src_sorted,
                        avg_num_neighbors if avg_num_neighbors > 0 else None, num_neigh)


@conv_fwd.register_fake
def _(x, sh, emb, mlp_weights, rowptr, perm, src_sorted, plan_id, avg_num_neighbors, num_neigh):
    return x.new_empty((x.shape[0], _plan(plan_id).plan.out_dim))


@torch.library.custom_op("matten_b200::conv_bwd", mutates_args=())
def conv_bwd(x: Tensor, sh: Tensor, emb: Tensor, mlp_weights: List[Tensor], rowptr: Tensor, perm: Tensor,
             src_sorted: Tensor, sender_ptr: Tensor, sender_perm: Tensor, plan_id: int, avg_num_neighbors: float,
             num_neigh: Optional[Tensor], grad_out: Tensor) -> List[Tensor]:
    """``mt_conv_bwd``: returns [grad_x, grad_w0, grad_w1, ...]."""
    gx, gws = ops.conv_bwd(_plan(plan_id), x, sh, emb, list(mlp_weights), rowptr, perm, src_sorted, sender_ptr,
                           sender_perm, avg_num_neighbors if avg_num_neighbors > 0 else None, num_neigh, grad_out)
    return [gx] + list(gws)


@conv_bwd.register_fake
def _(x, sh, emb, mlp_weights, rowptr, perm, src_sorted, sender_ptr, sender_perm, plan_id, avg_num_neighbors,
      num_neigh, grad_out):
    return [torch.empty_like(x)] + [torch.empty_like(w) for w in mlp_weights]


def _conv_setup(ctx, inputs, output):
    x, sh, emb, ws, rowptr, perm, src_sorted, plan_id, avg, num_neigh = inputs
    ctx.plan_id, ctx.avg, ctx.nw = plan_id, avg, len(ws)
    ctx.has_nn = num_neigh is not None
    ctx.save_for_backward(x, sh, emb, rowptr, perm, src_sorted, *( [num_neigh] if ctx.has_nn else []), *ws)


def _conv_bwd(ctx, g):
    saved = list(ctx.saved_tensors)
    x, sh, emb, rowptr, perm, src_sorted = saved[:6]
    num_neigh = saved[6] if ctx.has_nn else None
    ws = saved[6 + int(ctx.has_nn):]
    sptr, sperm = ops.csr_by_key(src_sorted.to(torch.int64), x.shape[0], True, None)
    outs = torch.ops.matten_b200.conv_bwd(x, sh, emb, ws, rowptr, perm, src_sorted, sptr, sperm, ctx.plan_id, ctx.avg,
                                          num_neigh, g.contiguous())
    return outs[0], None, None, list(outs[1:]), None, None, None, None, None, None


conv_fwd.register_autograd(_conv_bwd, setup_context=_conv_setup)


# ------------------------------------------------------------------------------------------ irreps linear
@torch.library.custom_op("matten_b200::linear_fwd", mutates_args=())
def linear_fwd(x: Tensor, weight: Tensor, species_perm: Optional[Tensor], species_ptr: Optional[Tensor],
               plan_id: int) -> Tensor:
    return ops.linear_fwd(_plan(plan_id), x, weight, species_perm, species_ptr)


@linear_fwd.register_fake
def _(x, weight, species_perm, species_ptr, plan_id):
    return x.new_empty(tuple(x.shape[:-1]) + (_plan(plan_id).out_dim,))


@torch.library.custom_op("matten_b200::linear_bwd", mutates_args=())
def linear_bwd(x: Tensor, weight: Tensor, grad_out: Tensor, species_perm: Optional[Tensor],
               species_ptr: Optional[Tensor], plan_id: int) -> Tuple[Tensor, Tensor]:
    gx, gw = ops.linear_bwd(_plan(plan_id), x, weight, grad_out, species_perm, species_ptr)
    return gx, gw.reshape(weight.shape)


@linear_bwd.register_fake
def _(x, weight, grad_out, species_perm, species_ptr, plan_id):
    return torch.empty_like(x), torch.empty_like(weight)


def _linear_setup(ctx, inputs, output):
    x, weight, sperm, sptr, plan_id = inputs
    ctx.plan_id, ctx.sp = plan_id, (sperm, sptr)
    ctx.save_for_backward(x, weight)


def _linear_bwd(ctx, g):
    x, weight = ctx.saved_tensors
    gx, gw = torch.ops.matten_b200.linear_bwd(x, weight, g.contiguous(), ctx.sp[0], ctx.sp[1], ctx.plan_id)
    return gx, gw, None, None, None


linear_fwd.register_autograd(_linear_bwd, setup_context=_linear_setup)


# ------------------------------------------------------------------------------------------ gate
@torch.library.custom_op("matten_b200::gate_fwd", mutates_args=())
def gate_fwd(x: Tensor, src_idx: Tensor, gate_idx: Tensor, act_id: Tensor, act_cst: Tensor, inv_first: Tensor,
             inv_count: Tensor) -> Tensor:
    """e3nn Gate as element tables (``plan.GatePlan``); the inverse tables are only used by the backward."""
    return ops.gate_fwd(x, x.shape[-1], src_idx.shape[0], src_idx, gate_idx, act_id, act_cst)


@gate_fwd.register_fake
def _(x, src_idx, gate_idx, act_id, act_cst, inv_first, inv_count):
    return x.new_empty(tuple(x.shape[:-1]) + (src_idx.shape[0],))


def _gate_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _gate_bwd(ctx, g):
    x, src, gidx, act, cst, inv_first, inv_count = ctx.saved_tensors
    gx = ops.gate_bwd(x, g.contiguous(), x.shape[-1], src.shape[0], src, gidx, act, cst, inv_first, inv_count)
    return gx, None, None, None, None, None, None


gate_fwd.register_autograd(_gate_bwd, setup_context=_gate_setup)
