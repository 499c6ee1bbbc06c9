"""Optimizer of the training harness: ``torch.optim.Adam`` / ``AdamW`` semantics
(the reference trains with ``torch.optim.Adam(lr=2e-4, weight_decay=1e-5)``, main.py:224-233) with the
whole parameter list updated by ONE CUDA launch (``rrt_adam_step``, csrc/optim.cu).

It is a ``torch.optim.Optimizer`` (param groups, ``zero_grad``, ``state_dict`` with the standard
``step`` / ``exp_avg`` / ``exp_avg_sq`` entries), so LR schedulers and checkpoints of the reference's
loop keep working.  CUDA fp32 parameters only; anything else raises (no fallback).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import cabi


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                                      decoupled=bool(decoupled)))

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = cabi.lib()
        for group in self.param_groups:
            by_step = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                    raise RuntimeError("rrt_mil_b200.optim.Adam updates float32 CUDA parameters only")
                if p.grad.is_sparse:
                    raise RuntimeError("sparse gradients are not supported")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] = int(st["step"]) + 1
                if not p.is_contiguous():
                    raise RuntimeError("parameters must be contiguous")
                by_step.setdefault((st["step"], p.device), []).append((p, p.grad.contiguous(), st))
            for (step, dev), items in by_step.items():
                arr = (cabi.RrtAdamTensor * len(items))()
                for a, (p, g, st) in zip(arr, items):
                    a.param, a.grad = p.data_ptr(), g.data_ptr()
                    a.exp_avg, a.exp_avg_sq, a.n = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()
                with torch.cuda.device(dev):
                    rc = lib.rrt_adam_step(arr, len(items), group["lr"], group["betas"][0], group["betas"][1],
                                           group["eps"], group["weight_decay"], int(group["decoupled"]), step,
                                           float(grad_scale), torch.cuda.current_stream(dev).cuda_stream)
                cabi.check(rc, "rrt_adam_step")
                # the kernel wrote the parameters through raw pointers: tell autograd (and every cache keyed
                # on the version counter, e.g. the encoder's fp16 weight shadows) that they changed
                for p, _, st in items:
                    torch.autograd.graph.increment_version(p)
                    torch.autograd.graph.increment_version(st["exp_avg"])
                    torch.autograd.graph.increment_version(st["exp_avg_sq"])
        return loss


class AdamW(Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, lr, betas, eps, weight_decay, decoupled=True)
