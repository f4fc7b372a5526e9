"""Lookahead wrapper: k fast steps, one slow interpolation (Zhang et al., 2019).

Reference: utils/lookahead.py:8-72 (vendored there from alphadl/lookahead.pytorch).
Behaviour preserved bit-for-bit:

* the wrapper *shares* the inner optimizer's ``param_groups`` (same list, same
  dicts), and stores a per-group ``counter``;
* ``step()`` = inner step, then for every group: if ``counter == 0`` do the slow
  update; ``counter += 1``; wrap to 0 at ``k``.  With k=10 the slow update fires
  on optimizer steps 1, 11, 21, ...;
* slow weights are created lazily as a copy of the fast weights *after* the
  first fast step, so the step-1 update is the identity;
* slow update: ``slow += alpha * (fast - slow); fast <- slow``; momentum buffers
  are left untouched.

Differences on purpose: this class is a real ``torch.optim.Optimizer`` (the
reference skipped ``Optimizer.__init__`` and breaks ``zero_grad`` /
``load_state_dict`` on torch 2.x), the slow state round-trips through
``state_dict`` by *parameter index* instead of ``id(tensor)``, and on CUDA the
interpolation is folded into the inner optimizer's single fused kernel.
"""
from __future__ import annotations

from typing import Any, Dict

import torch
from torch.optim import Optimizer


class Lookahead(Optimizer):
    def __init__(self, optimizer: Optimizer, k: int = 5, alpha: float = 0.5):
        if k < 1:
            raise ValueError(f"k must be >= 1, got {k}")
        if not 0.0 <= alpha <= 1.0:
            raise ValueError(f"alpha must be in [0, 1], got {alpha}")
        self.optimizer = optimizer
        self.k = int(k)
        self.alpha = float(alpha)
        super().__init__(optimizer.param_groups, {})
        self.param_groups = optimizer.param_groups  # one shared list, as in the reference
        self.fast_state = optimizer.state
        for group in self.param_groups:
            group.setdefault("counter", 0)

    # -- slow/fast interpolation -------------------------------------------
    @torch.no_grad()
    def update(self, group: Dict[str, Any]) -> None:
        fast, slow = [], []
        for p in group["params"]:
            st = self.state[p]
            if "slow_param" not in st:
                st["slow_param"] = p.detach().clone(memory_format=torch.preserve_format)
            fast.append(p)
            slow.append(st["slow_param"])
        if not fast:
            return
        delta = torch._foreach_sub(fast, slow)
        torch._foreach_add_(slow, delta, alpha=self.alpha)
        torch._foreach_copy_(fast, slow)

    def update_lookahead(self) -> None:
        for group in self.param_groups:
            self.update(group)

    # -- optimizer protocol ---------------------------------------------------
    def step(self, closure=None):
        fused = getattr(self.optimizer, "_fused", None)
        if fused is not None:
            # one kernel: SGD update + (device-side counter) Lookahead interpolation
            return self.optimizer.step(closure, _lookahead=self)
        loss = self.optimizer.step(closure)
        for group in self.param_groups:
            if group["counter"] == 0:
                self.update(group)
            group["counter"] = (group["counter"] + 1) % self.k
        return loss

    def zero_grad(self, set_to_none: bool = True) -> None:
        self.optimizer.zero_grad(set_to_none=set_to_none)

    def add_param_group(self, param_group: Dict[str, Any]) -> None:
        if getattr(self, "optimizer", None) is None or self.param_groups is not self.optimizer.param_groups:
            # called from Optimizer.__init__: adopt the inner optimizer's group dicts as-is
            self.param_groups.append(param_group)
            return
        param_group["counter"] = 0
        self.optimizer.add_param_group(param_group)  # shared list: visible here too

    # -- serialisation ----------------------------------------------------------
    def _param_index(self) -> Dict[int, int]:
        index, n = {}, 0
        for group in self.param_groups:
            for p in group["params"]:
                index[id(p)] = n
                n += 1
        return index

    def state_dict(self) -> Dict[str, Any]:
        fused = getattr(self.optimizer, "_fused", None)
        if fused is not None:
            fused.sync_to_python(self.optimizer, self)
        inner = self.optimizer.state_dict()
        index = self._param_index()
        slow = {index[id(p)]: {"slow_param": st["slow_param"]}
                for p, st in self.state.items() if "slow_param" in st}
        return {"fast_state": inner["state"], "slow_state": slow,
                "param_groups": inner["param_groups"], "k": self.k, "alpha": self.alpha}

    def load_state_dict(self, state_dict: Dict[str, Any]) -> None:
        self.optimizer.load_state_dict({"state": state_dict["fast_state"],
                                        "param_groups": state_dict["param_groups"]})
        self.param_groups = self.optimizer.param_groups
        self.fast_state = self.optimizer.state
        flat = [p for g in self.param_groups for p in g["params"]]
        self.state.clear()
        for idx, st in state_dict["slow_state"].items():
            p = flat[int(idx)]
            self.state[p]["slow_param"] = st["slow_param"].to(device=p.device, dtype=p.dtype).clone()
        self.k = int(state_dict.get("k", self.k))
        self.alpha = float(state_dict.get("alpha", self.alpha))
        for group in self.param_groups:
            group.setdefault("counter", 0)
        fused = getattr(self.optimizer, "_fused", None)
        if fused is not None:
            fused.sync_from_python(self.optimizer, self)
