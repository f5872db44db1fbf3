"""Step tail: flat parameter/gradient buffers, fused Adam, Noam schedule.

FlatParams re-homes every parameter of a module into ONE contiguous fp32 buffer (and its gradient into another),
so the data-parallel step is a single NCCL all-reduce and the optimizer a single kernel launch.
"""
from __future__ import annotations

import math

import torch

from . import _lib as L
from . import ops


class FlatParams:
    """Views of all trainable parameters (and their .grad) into two flat fp32 buffers.

    The gradient buffer carries `extra` trailing floats (used for [sum-loss, n_tokens]) so that they ride in
    the same all-reduce as the gradients."""

    def __init__(self, module: torch.nn.Module, extra: int = 2, tail=("conv",)):
        """tail: names of submodules whose parameters go to the END of the buffers (default: the CNN front end `conv`).  Their
        gradients are the last ones backward produces, so everything before `tail_offset` can be all-reduced while the front
        end's backward is still running (parallel.DataParallelStep)."""
        seen, params = set(), []
        tail_ids = set()
        for name in tail:
            sub = getattr(module, name, None)
            if isinstance(sub, torch.nn.Module):
                tail_ids.update(id(p) for p in sub.parameters())

        def take(p, tail_pass=False):
            if p is not None and p.requires_grad and id(p) not in seen and (tail_pass or id(p) not in tail_ids):
                seen.add(id(p))
                params.append(p)

        # Q/K/V projection weights (then their biases) of every attention block go back to back, so that the block can run
        # them -- and their gradients -- as one stacked GEMM operand (ops.AttnProjFn); everything else in module order
        for m in module.modules():
            lin = [getattr(m, n, None) for n in ("query_linear", "key_linear", "value_linear")]
            if all(isinstance(l, torch.nn.Linear) for l in lin):
                for l in lin:
                    take(l.weight)
                for l in lin:
                    take(l.bias)
        for p in module.parameters():
            take(p)
        n_head = len(params)
        for p in module.parameters():
            take(p, tail_pass=True)
        self.params = params
        dev = params[0].device
        sizes = [p.numel() for p in params]
        # 16-byte aligned slots so every view keeps vector-load alignment
        self.offsets, off = [], 0
        for n in sizes:
            self.offsets.append(off)
            off += (n + 3) // 4 * 4
        self.numel = off
        self.tail_offset = self.offsets[n_head] if n_head < len(params) else off      # first element of the tail parameters
        self.extra = extra
        self.flat = torch.zeros(off, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(off + extra, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, o, n in zip(params, self.offsets, sizes):
                self.flat[o:o + n].copy_(p.detach().reshape(-1))
                p.data = self.flat[o:o + n].view(p.shape)
                p.grad = self.flat_grad[o:o + n].view(p.shape)
                p._b200_flat_grad = True      # opt-in for ops._grad_sink: backward kernels may accumulate straight into .grad
        if dev.type == "cuda":
            ops.weight_cache = ops.WeightOperandCache(self)     # bf16 operands of all weights: one conversion launch per step

    @property
    def extras(self) -> torch.Tensor:
        return self.flat_grad[self.numel:]

    def zero_grad(self):
        self.flat_grad.zero_()

    def ensure_grad_views(self):
        """Re-attach .grad views if something (e.g. zero_grad(set_to_none=True)) dropped them."""
        for p, o in zip(self.params, self.offsets):
            n = p.numel()
            view = self.flat_grad[o:o + n].view(p.shape)
            if p.grad is None:
                p.grad = view
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
                p.grad = view


class FusedAdam:
    """torch.optim.Adam(betas, eps) semantics (utils/functions.py:107) as one kernel over FlatParams."""

    def __init__(self, flat: FlatParams, lr=1e-3, betas=(0.9, 0.98), eps=1e-9):
        self.flat = flat
        self.param_groups = [dict(lr=lr, betas=betas, eps=eps, params=flat.params)]
        self.m = torch.zeros_like(flat.flat)
        self.v = torch.zeros_like(flat.flat)
        self.t = 0

    def zero_grad(self, set_to_none: bool = False):
        self.flat.zero_grad()

    def step(self, grad_scale: float = 1.0, grad_scale_dev: torch.Tensor = None):
        g = self.param_groups[0]
        self.t += 1
        f = self.flat
        L.check(L.load().b200asr_adam_step(L.ptr(f.flat), L.ptr(f.flat_grad), L.ptr(self.m), L.ptr(self.v), f.numel,
                                           float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                                           self.t, float(grad_scale), L.ptr(grad_scale_dev),
                                           torch.cuda.current_stream().cuda_stream), "adam_step")
        if ops.weight_cache is not None and ops.weight_cache.flat is f:
            ops.weight_cache.after_optimizer_step()

    def grad_sumsq(self) -> torch.Tensor:
        out = torch.zeros(1, device=self.flat.flat.device, dtype=torch.float32)
        L.check(L.load().b200asr_sumsq(L.ptr(self.flat.flat_grad), self.flat.numel, L.ptr(out),
                                       torch.cuda.current_stream().cuda_stream), "sumsq")
        return out

    def state_dict(self):
        return dict(m=self.m, v=self.v, t=self.t, param_groups=[{k: v for k, v in self.param_groups[0].items() if k != "params"}])

    def load_state_dict(self, sd):
        self.m.copy_(sd["m"]); self.v.copy_(sd["v"]); self.t = int(sd["t"])


class NoamOpt:
    """utils/optimizer.py:3-32 -- same attributes and methods; `optimizer` may be FusedAdam or any torch optimizer."""

    def __init__(self, model_size, factor, warmup, optimizer, min_lr=1e-5):
        self.optimizer = optimizer
        self._step = 0
        self.warmup = warmup
        self.factor = factor
        self.model_size = model_size
        self._rate = 0
        self.min_lr = min_lr

    def step(self, **kw):
        self._step += 1
        rate = self.rate()
        for g in self.optimizer.param_groups:
            g["lr"] = rate
        self._rate = rate
        self.optimizer.step(**kw)

    def zero_grad(self):
        self.optimizer.zero_grad()

    def rate(self, step=None):
        step = self._step
        return max(self.min_lr, self.factor * (self.model_size ** (-0.5) * min(step ** (-0.5), step * self.warmup ** (-1.5))))
