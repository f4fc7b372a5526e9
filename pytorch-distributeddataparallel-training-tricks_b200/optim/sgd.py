"""SGD with momentum / Nesterov (reference call site: utils/train.py:41).

Update rule (dampening 0, weight decay optional), identical to ``torch.optim.SGD``::

    g   = grad + wd * p
    buf = g                       (first step)      |  buf = mu * buf + g
    d   = g + mu * buf            (nesterov)        |  d   = buf
    p  -= lr * d

Two execution paths share this class:

* **flat / fused (CUDA)** -- parameters, gradients, momentum (and Lookahead slow
  weights) live in contiguous fp32 arenas (see :mod:`ddpt_b200.parallel.arena`)
  and one sm_100a kernel (csrc/optim/fused_lookahead_sgd.cu) applies the whole
  update, including the optional Lookahead interpolation, the found-inf skip
  predicate and the bf16 re-cast of the weights for the tensor-core path.  The
  learning rate is read from device memory so a captured CUDA graph stays valid
  when schedulers change it.
* **eager (CPU / fallback)** -- ``torch._foreach`` ops, used by the gloo plumbing
  configuration and as the oracle in tests.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
from torch.optim import Optimizer


class SGD(Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, momentum: float = 0.0,
                 dampening: float = 0.0, weight_decay: float = 0.0, nesterov: bool = False):
        if lr < 0.0 or momentum < 0.0 or weight_decay < 0.0:
            raise ValueError("lr, momentum and weight_decay must be non-negative")
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening,
                                      weight_decay=weight_decay, nesterov=nesterov))
        self._fused = None  # FusedState, attached by ops.fused_optim.attach()

    # ------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None, *, _lookahead=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._fused is not None:
            self._fused.step(self, _lookahead)
            return loss
        for group in self.param_groups:
            self._eager_group_step(group)
        return loss

    def _eager_group_step(self, group: dict) -> None:
        lr, mu, damp, wd, nesterov = (group["lr"], group["momentum"], group["dampening"],
                                      group["weight_decay"], group["nesterov"])
        params, grads, bufs, fresh = [], [], [], []
        for p in group["params"]:
            if p.grad is None:
                continue
            if p.grad.is_sparse:
                raise RuntimeError("SGD does not support sparse gradients")
            params.append(p)
            grads.append(p.grad)
            state = self.state[p]
            if mu != 0 and "momentum_buffer" not in state:
                state["momentum_buffer"] = None
            bufs.append(state.get("momentum_buffer"))
        if not params:
            return
        if wd != 0:
            grads = torch._foreach_add(grads, params, alpha=wd)
        if mu != 0:
            for i, p in enumerate(params):
                if bufs[i] is None:
                    bufs[i] = self.state[p]["momentum_buffer"] = torch.clone(grads[i]).detach()
                    fresh.append(i)
            old = [i for i in range(len(params)) if i not in fresh]
            if old:
                ob = [bufs[i] for i in old]
                torch._foreach_mul_(ob, mu)
                torch._foreach_add_(ob, [grads[i] for i in old], alpha=1 - damp)
            grads = torch._foreach_add(grads, bufs, alpha=mu) if nesterov else bufs
        torch._foreach_add_(params, grads, alpha=-lr)

    def zero_grad(self, set_to_none: bool = True) -> None:
        if self._fused is not None:
            self._fused.zero_grad(set_to_none)
            return
        super().zero_grad(set_to_none=set_to_none)
