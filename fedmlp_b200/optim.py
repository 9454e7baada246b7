"""Fused Adam on a flat client model (kernel adam.cu).

`FlatAdam(module)` flattens the module's parameters / buffers into one fp32 buffer
(`flat.flatten_module_`), keeps the gradients and both Adam moments in flat buffers of the same
layout, and performs `step()` (+ optional `zero_grad`) as ONE kernel launch instead of ~10 foreach
kernels over 364 tensors.  Same hyper-parameters and arithmetic as the optimizer the reference
creates every round: torch.optim.Adam(net.parameters(), lr, betas=(0.9, 0.999), weight_decay=5e-4)
(utils/local_training.py:912-913, :1149-1150).  One difference from torch.optim.Adam: a trainable parameter
that received NO gradient in a step is updated with a zero gradient (decay + moment decay) instead of being
skipped; frozen parameters (requires_grad=False) are left alone like in torch.
"""
from __future__ import annotations

import torch

from . import _cabi as cabi
from .flat import FlatStateDict, flatten_module_

_CHUNK = cabi.FEDAVG_CHUNK


class FlatAdam:
    def __init__(self, module: torch.nn.Module, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                 flat: FlatStateDict | None = None):
        self.module = module
        self.flat = flat if flat is not None else flatten_module_(module)
        buf = self.flat.flat_f32
        cabi.require_cuda(buf)
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.grad = torch.zeros_like(buf)
        self.exp_avg = torch.zeros_like(buf)
        self.exp_avg_sq = torch.zeros_like(buf)
        self.step_count = 0
        lay = self.flat.layout
        offset_of = {k: (lay.offsets[i], lay.numels[i]) for i, k in enumerate(lay.keys) if not lay.is_int[i]}
        starts, lens = [], []
        self._bound = []          # (parameter, its gradient view): re-checked in step()
        for name, p in module.named_parameters():
            off, n = offset_of[name]
            if p.data_ptr() != buf.data_ptr() + 4 * off:
                raise ValueError(f"parameter {name} is not a view of the flat buffer")
            if not p.requires_grad:
                continue          # torch.optim.Adam never touches frozen parameters (no decay, no moments)
            g = self.grad[off:off + n].view_as(p)
            p.grad = g                                       # autograd accumulates in place from now on
            self._bound.append((p, g))
            for s in range(0, n, _CHUNK):
                starts.append(off + s)
                lens.append(min(_CHUNK, n - s))
        dev = buf.device
        self.chunk_start = torch.tensor(starts, dtype=torch.int64, device=dev)
        self.chunk_len = torch.tensor(lens, dtype=torch.int32, device=dev)
        self.n_chunks = len(starts)

    def reset(self, lr=None):
        """Fresh optimizer state (the reference re-creates Adam every round, :912, :1149): zero the moments and
        the step count, keep the buffers."""
        self.exp_avg.zero_(); self.exp_avg_sq.zero_(); self.grad.zero_()
        self.step_count = 0
        if lr is not None:
            self.lr = float(lr)
        self._rebind()

    def zero_grad(self, set_to_none=False):
        """Clears the flat gradient buffer.  (module.zero_grad() / set_to_none=True would detach the parameters'
        .grad views; step() re-binds them if that happened.)"""
        self.grad.zero_()
        self._rebind()

    def _rebind(self):
        for p, g in self._bound:
            if p.grad is None:
                p.grad = g                                   # detached by a zero_grad(set_to_none=True): nothing was accumulated
            elif p.grad.data_ptr() != g.data_ptr():
                g.copy_(p.grad)                              # autograd allocated a fresh .grad: take its values, alias again
                p.grad = g

    def step(self, zero_grad=False):
        """One Adam step; zero_grad=True also clears the gradients in the same launch."""
        self._rebind()
        self.step_count += 1
        buf = self.flat.flat_f32
        with torch.cuda.device(buf.device):
            cabi.check(cabi.lib().fmlp_adam_step_f32(
                buf.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                self.chunk_start.data_ptr(), self.chunk_len.data_ptr(), self.n_chunks, self.lr, self.betas[0],
                self.betas[1], self.eps, self.weight_decay, self.step_count, 1 if zero_grad else 0,
                cabi.stream_ptr(buf.device)), "fmlp_adam_step_f32")
