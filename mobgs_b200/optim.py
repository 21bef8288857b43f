"""Fused multi-tensor Adam behind torch.optim.Adam's interface.

The reference creates `torch.optim.Adam(l, lr=0.0, eps=1e-15)` over 17 named parameter groups per
Gaussian model (scene/gaussian_model.py:598-641) and its densification code edits
`optimizer.state[p]["exp_avg" | "exp_avg_sq"]` and `optimizer.param_groups` in place
(gaussian_model.py:1044-1123).  `FusedAdam` therefore *is* a torch.optim.Adam — same constructor,
same state layout, same param_groups — whose `step()` hands every tensor that has a gradient to one
`mobgs_adam_step` launch; `fused_step([...])` steps several optimisers (static + dynamic model) in the
same launch.  Only what the reference uses is supported: fp32 CUDA parameters, amsgrad=False,
weight_decay=0, maximize=False.  There is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib


def _collect(opt, items):
    for group in opt.param_groups:
        if group.get("amsgrad") or group.get("weight_decay", 0) != 0 or group.get("maximize"):
            raise RuntimeError("FusedAdam supports amsgrad=False, weight_decay=0, maximize=False only")
        beta1, beta2 = group["betas"]
        for p in group["params"]:
            if p.grad is None:
                continue
            if not (p.is_cuda and p.dtype == torch.float32 and p.grad.dtype == torch.float32):
                raise RuntimeError("FusedAdam needs fp32 CUDA parameters and gradients (no CPU fallback)")
            if p.grad.is_sparse:
                raise RuntimeError("FusedAdam does not support sparse gradients")
            st = opt.state[p]
            if len(st) == 0:      # same lazy state init as torch.optim.Adam._init_group
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["step"] += 1
            step = float(st["step"])
            if not p.is_contiguous():
                raise RuntimeError("FusedAdam needs contiguous parameters")
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            m, v = st["exp_avg"], st["exp_avg_sq"]
            if not (m.is_contiguous() and v.is_contiguous()):
                raise RuntimeError("FusedAdam needs contiguous optimiser state")
            items.append((p, g, m, v, group["lr"] / (1 - beta1 ** step), math.sqrt(1 - beta2 ** step),
                          beta1, beta2, group["eps"]))


def _launch(items):
    if not items:
        return
    chunk = _lib.load().mobgs_adam_chunk_elems()
    stream = _lib.current_stream()
    # one launch per (beta1, beta2, eps) combination (the reference has exactly one)
    by_hyper = {}
    for it in items:
        by_hyper.setdefault(it[6:], []).append(it)
    for (beta1, beta2, eps), its in by_hyper.items():
        for s in range(0, len(its), _lib.ADAM_MAX_TENSORS):
            part = its[s:s + _lib.ADAM_MAX_TENSORS]
            a = _lib.Adam()
            a.n_tensors = len(part)
            a.beta1, a.beta2, a.eps = beta1, beta2, eps
            a.one_minus_beta1, a.one_minus_beta2 = 1 - beta1, 1 - beta2
            chunks = 0
            for i, (p, g, m, v, step_size, bc2_sqrt, *_rest) in enumerate(part):
                a.param[i], a.grad[i] = p.data_ptr(), g.data_ptr()
                a.exp_avg[i], a.exp_avg_sq[i] = m.data_ptr(), v.data_ptr()
                a.numel[i] = p.numel()
                a.step_size[i], a.bc2_sqrt[i] = step_size, bc2_sqrt
                a.chunk_begin[i] = chunks
                chunks += (p.numel() + chunk - 1) // chunk
            a.chunk_begin[len(part)] = chunks
            _lib.call("mobgs_adam_step", a, stream)
    # The kernel wrote the parameters and the Adam moments through raw pointers: tell autograd (saved-tensor
    # checks) and every version-keyed cache (deformation._operand_pack) that they changed, as an
    # in-place torch op would have.
    torch.autograd.graph.increment_version([t for it in items for t in it[0:1] + it[2:4]])


@torch.no_grad()
def fused_step(optimizers) -> None:
    """Step several FusedAdam / torch.optim.Adam instances with one kernel launch (the static and the
    dynamic Gaussian model of train.py:796-800)."""
    items = []
    for opt in optimizers:
        _collect(opt, items)
    _launch(items)


class FusedAdam(torch.optim.Adam):
    """Drop-in for the reference's `torch.optim.Adam(l, lr=0.0, eps=1e-15)`."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, **kw):
        for k in ("amsgrad", "maximize", "capturable", "differentiable", "fused"):
            if kw.get(k):
                raise RuntimeError(f"FusedAdam: {k} is not supported")
        if kw.get("weight_decay", 0) != 0:
            raise RuntimeError("FusedAdam: weight_decay is not supported")
        super().__init__(params, lr=lr, betas=betas, eps=eps, foreach=False, **{k: v for k, v in kw.items() if k != "foreach"})

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        items = []
        _collect(self, items)
        _launch(items)
        return loss
