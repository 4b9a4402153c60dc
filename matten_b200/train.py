"""
Training step of the tensor-property models: forward -> MSE -> backward -> (NCCL all-reduce) -> Adam, all on
hand-written kernels (reference: ``BaseModel.shared_step`` / ``compute_loss`` in src/matten/model/model.py:234-274,
325-360 with ``MSELoss`` and ``torch.optim.Adam(lr, weight_decay)`` of scripts/configs/materials_tensor.yaml:103-107).

Parameters and gradients live in ONE flat buffer each (``param.data`` / ``param.grad`` are views), so that a step
is a single fused Adam launch and, under data parallelism, a single ``all_reduce`` of the flat gradient over
NCCL / NVLink.  BatchNorm statistics stay per rank, like Lightning + e3nn.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import autograd as A
from . import ops


class FlatAdam:
    """torch.optim.Adam (amsgrad off, L2 weight decay) over a flattened parameter set."""

    def __init__(self, params, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dt, dev = self.params[0].dtype, self.params[0].device
        if any(p.dtype != dt or p.device != dev for p in self.params):
            raise ValueError("all parameters must share dtype and device")
        if not self.params[0].is_cuda:
            raise RuntimeError("matten_b200 trains on CUDA only (there is no CPU fallback)")
        n = sum(p.numel() for p in self.params)
        self.flat_p = torch.empty(n, dtype=dt, device=dev)
        self.flat_g = torch.zeros(n, dtype=dt, device=dev)
        self.m = torch.zeros(n, dtype=dt, device=dev)
        self.v = torch.zeros(n, dtype=dt, device=dev)
        off = 0
        for p in self.params:
            k = p.numel()
            self.flat_p[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + k].view(p.shape)
            p.grad = self.flat_g[off:off + k].view(p.shape)
            off += k
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.t = 0

    def zero_grad(self):
        self.flat_g.zero_()

    def step(self, grad_scale: float = 1.0):
        self.t += 1
        ops.adam_step(self.flat_p, self.flat_g, self.m, self.v, self.t, self.lr, self.betas[0], self.betas[1],
                      self.eps, self.weight_decay, grad_scale)


class Trainer:
    """One data-parallel replica: ``step(batch, target)`` runs forward, loss, backward, gradient all-reduce (when
    ``torch.distributed`` is initialised with world size > 1) and the Adam update."""

    def __init__(self, model: torch.nn.Module, lr=1e-2, weight_decay=1e-5, betas=(0.9, 0.999), eps=1e-8,
                 process_group=None, output_key: Optional[str] = None):
        self.model = model
        self.opt = FlatAdam(model.parameters(), lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.pg = process_group
        self.output_key = output_key
        self.world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)

    def forward_loss(self, batch: Dict[str, torch.Tensor], target: torch.Tensor, atom_selector=None):
        out = self.model(batch)
        if isinstance(out, dict):
            key = self.output_key or next(iter(out))
            out = out[key]
        if atom_selector is not None:  # reference model.py:342-344
            out = out[atom_selector]
        return A.mse_loss(out, target), out

    def step(self, batch: Dict[str, torch.Tensor], target: torch.Tensor, atom_selector=None) -> torch.Tensor:
        self.model.train()
        self.opt.zero_grad()
        loss, _ = self.forward_loss(batch, target, atom_selector)
        loss.backward()
        if self.world > 1:
            torch.distributed.all_reduce(self.opt.flat_g, group=self.pg)
        self.opt.step(grad_scale=1.0 / self.world)
        return loss.detach()


    @torch.no_grad()
    def evaluate(self, batches) -> Dict[str, float]:
        """Validation pass (reference ``validation_step`` / metrics, model/model.py:300-360): mean squared and mean
        absolute error over all target components; ``batches`` yields (graph batch, target, atom selector | None)."""
        self.model.eval()
        se = torch.zeros((), dtype=torch.float64, device=self.opt.flat_p.device)
        ae = torch.zeros_like(se)
        n = 0
        for batch, target, sel in batches:
            out = self.model(batch)
            if isinstance(out, dict):
                out = out[self.output_key or next(iter(out))]
            if sel is not None:
                out = out[sel]
            d = (out - target).double()
            se += (d * d).sum()
            ae += d.abs().sum()
            n += d.numel()
        if self.world > 1:
            t = torch.stack([se, ae, torch.tensor(float(n), dtype=torch.float64, device=se.device)])
            torch.distributed.all_reduce(t, group=self.pg)
            se, ae, n = t[0], t[1], float(t[2])
        n = max(float(n), 1.0)
        return {"mse": float(se) / n, "mae": float(ae) / n}
