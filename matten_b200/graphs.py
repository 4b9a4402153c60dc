"""
CUDA-graph replay of the model forward (inference).

One forward of the lmax-2 model is ~55 kernel launches of 5 us .. 1.7 ms; issued one by one from Python the GPU idles
between the short ones.  ``CapturedForward`` records the whole forward -- index bookkeeping (radix sort, CSR), edge
embedding, the fused convolutions, linears, gates, pooling -- ONCE per batch shape into a CUDA graph and replays it;
new batches of the same shape are copied into the graph's static input buffers.  Nothing is traced or compiled: the
graph holds exactly the kernels the eager path launches.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from . import ops


class CapturedForward:
    def __init__(self, model: torch.nn.Module, warmup: int = 2):
        self.model = model
        self.warmup = warmup
        self._graphs: Dict[Tuple, tuple] = {}
        self._tensors = None

    def _model_version(self) -> int:
        """Changes whenever a parameter or buffer is written in place (optimizer step, load_state_dict): derived
        vectors cached outside the graph (BatchNorm.eval_affine) would otherwise go stale inside a capture."""
        ts = self._tensors
        if ts is None:  # walking the module tree costs ~0.3 ms: do it once (call invalidate() after model.to(...))
            ts = self._tensors = list(self.model.parameters()) + list(self.model.buffers())
        return hash(tuple(t._version for t in ts))

    def invalidate(self):
        """Forget the captures and the cached tensor list (after the model's tensors were replaced, e.g. ``.to``)."""
        self._tensors = None
        self._graphs.clear()

    def _signature(self, batch, slot: int = 0) -> Tuple:
        return (slot, self._model_version()) + tuple(sorted((k, tuple(v.shape), str(v.dtype)) if isinstance(v, torch.Tensor) else (k, v)
                                      for k, v in batch.items() if not k.startswith("_")))

    def _capture(self, batch):
        dev = next(self.model.parameters()).device
        static = {k: (v.to(dev).clone() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()
                  if not k.startswith("_")}
        self.model.eval()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(self.warmup):  # lazy initialisation (plan tables, function attributes) outside the capture
                self.model(dict(static), check=False)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(g):
            out = self.model(dict(static), check=False)
            last = getattr(self.model, "_last_graph", None)
            flag = last.flag if last is not None else None  # zeroed by a captured fill at every replay
        return g, static, out, flag

    def __call__(self, batch, check: bool = True, slot: int = 0):
        """Same contract as ``model(batch)``: returns the dict of predictions (tensors owned by the graph: valid until
        the next call with the same shape and slot).  ``slot`` selects one of several independent captures of the same
        shape (double buffering: the inputs of step i+1 are copied in while step i runs)."""
        sig = self._signature(batch, slot)
        ent = self._graphs.get(sig)
        if ent is None:
            ent = self._capture(batch)
            self._graphs[sig] = ent
        g, static, out, flag = ent
        for k, v in batch.items():
            if isinstance(v, torch.Tensor) and k in static and v.data_ptr() != static[k].data_ptr():
                static[k].copy_(v, non_blocking=True)
        g.replay()
        if check and flag is not None:
            ops.raise_on_flag(flag)
        return out

    def static_inputs(self, batch, slot: int = 0):
        """The graph's input buffers for this batch shape (capture on first use); writing into them directly (e.g. a
        pinned-host -> device copy) saves the extra device-to-device copy of ``__call__``."""
        sig = self._signature(batch, slot)
        if sig not in self._graphs:
            self._graphs[sig] = self._capture(batch)
        return self._graphs[sig][1]

    def error_flag(self, batch, slot: int = 0):
        """Device error word of the capture for this batch shape / slot (read it after the replay has finished)."""
        return self._graphs[self._signature(batch, slot)][3]
