"""Learning-rate schedules used by the reference driver (utils/train.py:48-53,104-106).

``LambdaLR`` and ``ReduceLROnPlateau`` follow the torch semantics the reference
relied on, including the two behaviours that shape its logged LR curve:

* constructing ``LambdaLR`` performs an initial step, so epoch 0 trains at
  ``base_lr * lambda(0)`` = 0 for the warm-up lambdas (quirk Q3);
* the driver only steps the warm-up scheduler while ``epoch <= warmup_epochs``,
  so the cosine branch is evaluated exactly once, at ``lambda(W+1)`` (quirk Q2).

They mutate ``optimizer.param_groups[i]["lr"]`` (a Python float); the fused
optimizer mirrors that value into device memory lazily, so schedules never force
a CUDA-graph re-capture.
"""
from __future__ import annotations

import math
from typing import Callable, List, Sequence, Union


def warmup_lambda(kind: str, warmup_epochs: int, total_epochs: int) -> Callable[[int], float]:
    """lr multiplier: ramp ``e/W`` for ``e <= W``; then 1 (linear) or half-cosine to 0 (cosine)."""
    w, t = warmup_epochs, total_epochs
    if kind == "linear":
        return lambda e: e / w if e <= w else 1
    if kind == "cosine":
        return lambda e: e / w if e <= w else 0.5 * (math.cos((e - w) / (t - w) * math.pi) + 1)
    raise ValueError(f"warmup kind must be 'linear' or 'cosine', got {kind!r}")


class LambdaLR:
    def __init__(self, optimizer, lr_lambda: Union[Callable[[int], float], Sequence[Callable[[int], float]]]):
        self.optimizer = optimizer
        groups = optimizer.param_groups
        self.lr_lambdas: List[Callable[[int], float]] = (
            list(lr_lambda) if isinstance(lr_lambda, (list, tuple)) else [lr_lambda] * len(groups))
        if len(self.lr_lambdas) != len(groups):
            raise ValueError("need one lr_lambda per param group")
        for g in groups:
            g.setdefault("initial_lr", g["lr"])
        self.base_lrs = [g["initial_lr"] for g in groups]
        self.last_epoch = -1
        self.step()  # torch semantics: lr is set to base_lr * lambda(0) on construction

    def step(self) -> None:
        self.last_epoch += 1
        for g, base, fn in zip(self.optimizer.param_groups, self.base_lrs, self.lr_lambdas):
            g["lr"] = base * fn(self.last_epoch)

    def get_last_lr(self) -> List[float]:
        return [g["lr"] for g in self.optimizer.param_groups]

    def state_dict(self) -> dict:
        return {"last_epoch": self.last_epoch, "base_lrs": list(self.base_lrs)}

    def load_state_dict(self, state: dict) -> None:
        self.last_epoch = int(state["last_epoch"])
        self.base_lrs = list(state["base_lrs"])


class ReduceLROnPlateau:
    def __init__(self, optimizer, mode: str = "min", factor: float = 0.1, patience: int = 10,
                 threshold: float = 1e-4, threshold_mode: str = "rel", cooldown: int = 0,
                 min_lr: float = 0.0, eps: float = 1e-8, verbose: bool = False):
        if factor >= 1.0:
            raise ValueError("factor should be < 1.0")
        if mode not in ("min", "max") or threshold_mode not in ("rel", "abs"):
            raise ValueError("bad mode / threshold_mode")
        self.optimizer = optimizer
        self.mode, self.factor, self.patience = mode, factor, patience
        self.threshold, self.threshold_mode = threshold, threshold_mode
        self.cooldown, self.eps, self.verbose = cooldown, eps, verbose
        self.min_lrs = [min_lr] * len(optimizer.param_groups)
        self.best = math.inf if mode == "min" else -math.inf
        self.num_bad_epochs = 0
        self.cooldown_counter = 0
        self.last_epoch = 0

    def _is_better(self, value: float) -> bool:
        if self.mode == "min":
            bound = self.best * (1.0 - self.threshold) if self.threshold_mode == "rel" else self.best - self.threshold
            return value < bound
        bound = self.best * (1.0 + self.threshold) if self.threshold_mode == "rel" else self.best + self.threshold
        return value > bound

    def step(self, metrics: float) -> None:
        current = float(metrics)
        self.last_epoch += 1
        if self._is_better(current):
            self.best, self.num_bad_epochs = current, 0
        else:
            self.num_bad_epochs += 1
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.num_bad_epochs = 0
        if self.num_bad_epochs > self.patience:
            self._reduce(self.last_epoch)
            self.cooldown_counter = self.cooldown
            self.num_bad_epochs = 0

    def _reduce(self, epoch: int) -> None:
        for i, g in enumerate(self.optimizer.param_groups):
            old = float(g["lr"])
            new = max(old * self.factor, self.min_lrs[i])
            if old - new > self.eps:
                g["lr"] = new
                if self.verbose:
                    print(f"Epoch {epoch:5d}: reducing learning rate of group {i} to {new:.4e}.")

    def state_dict(self) -> dict:
        return {k: getattr(self, k) for k in ("best", "num_bad_epochs", "cooldown_counter", "last_epoch")}

    def load_state_dict(self, state: dict) -> None:
        for k, v in state.items():
            setattr(self, k, v)
