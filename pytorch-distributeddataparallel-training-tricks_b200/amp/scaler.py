"""Dynamic loss scaling with Apex's constants.

scale starts at 2**16; an overflow halves it and skips that optimizer step; 2000
consecutive clean steps double it, capped at 2**24.

State lives where the gradients live.  On CUDA the scale, its reciprocal, the
overflow flag and the clean-step counter are five words of device memory that
the fused all-reduce / unscale kernel, the fused optimizer and the scale-update
kernel read and write directly -- the host never looks at them inside the step,
so a captured CUDA graph replays the whole protocol.  On CPU the same numbers
are Python scalars.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch


class LossScaler:
    INIT_SCALE = 2.0 ** 16
    MAX_SCALE = 2.0 ** 24
    MIN_SCALE = 1.0
    GROWTH_INTERVAL = 2000

    def __init__(self, device: torch.device, dynamic: bool = True, init_scale: Optional[float] = None):
        self.device = torch.device(device)
        self.dynamic = dynamic
        scale = float(init_scale if init_scale is not None else (self.INIT_SCALE if dynamic else 1.0))
        self.on_device = self.device.type == "cuda"
        if self.on_device:
            # [scale, 1/scale, found_inf, clean_steps, skipped_total]
            self.state = torch.tensor([scale, 1.0 / scale, 0.0, 0.0, 0.0], dtype=torch.float32,
                                      device=self.device)
        else:
            self._scale, self._clean, self._found_inf, self._skipped = scale, 0, False, 0

    # -- views used by kernels ---------------------------------------------------
    @property
    def scale_tensor(self) -> torch.Tensor:
        return self.state[0:1]

    @property
    def inv_scale_tensor(self) -> torch.Tensor:
        return self.state[1:2]

    @property
    def found_inf_tensor(self) -> torch.Tensor:
        return self.state[2:3]

    # -- host accessors (sync on CUDA; never called inside the step) ------------------
    def loss_scale(self) -> float:
        return float(self.state[0].item()) if self.on_device else self._scale

    def skipped_steps(self) -> int:
        return int(self.state[4].item()) if self.on_device else self._skipped

    # -- protocol --------------------------------------------------------------------
    def scale(self, loss: torch.Tensor) -> torch.Tensor:
        if self.on_device:
            return loss.float() * self.state[0]
        return loss.float() * self._scale

    def begin_step(self) -> None:
        """Clear the overflow flag before gradients are produced."""
        if self.on_device:
            self.state[2:3].zero_()
        else:
            self._found_inf = False

    def unscale_(self, grads: Iterable[torch.Tensor], already_unscaled: bool = False) -> None:
        """grads *= 1/scale and record overflow.  ``already_unscaled``: the fused
        all-reduce did both in the same pass (only the flag is consumed here)."""
        if already_unscaled:
            return
        grads = [g for g in grads if g is not None]
        if not grads:
            return
        if self.on_device:
            from ..ops import multi_tensor
            multi_tensor.unscale_and_check(grads, self.state)
            return
        inv = 1.0 / self._scale
        bad = False
        for g in grads:
            if inv != 1.0:
                g.mul_(inv)
            if not bool(torch.isfinite(g).all()):
                bad = True
        self._found_inf = self._found_inf or bad

    def update(self) -> None:
        """Post-step bookkeeping (halve on overflow / grow after a clean interval)."""
        if not self.dynamic:
            return
        if self.on_device:
            from ..ops import multi_tensor
            multi_tensor.update_scale(self.state, self.GROWTH_INTERVAL, self.MIN_SCALE, self.MAX_SCALE)
            return
        if self._found_inf:
            self._scale = max(self._scale / 2.0, self.MIN_SCALE)
            self._clean = 0
            self._skipped += 1
        else:
            self._clean += 1
            if self._clean >= self.GROWTH_INTERVAL:
                self._scale = min(self._scale * 2.0, self.MAX_SCALE)
                self._clean = 0

    def found_inf_host(self) -> bool:
        return bool(self.state[2].item() != 0) if self.on_device else self._found_inf

    # -- serialisation -------------------------------------------------------------------
    def state_dict(self) -> dict:
        if self.on_device:
            s = self.state.tolist()
            return {"scale": s[0], "clean_steps": int(s[3]), "skipped": int(s[4]), "dynamic": self.dynamic}
        return {"scale": self._scale, "clean_steps": self._clean, "skipped": self._skipped,
                "dynamic": self.dynamic}

    def load_state_dict(self, state: dict) -> None:
        self.dynamic = bool(state.get("dynamic", self.dynamic))
        scale = float(state["scale"])
        if self.on_device:
            self.state.copy_(torch.tensor([scale, 1.0 / scale, 0.0, float(state["clean_steps"]),
                                           float(state.get("skipped", 0))]))
        else:
            self._scale, self._clean = scale, int(state["clean_steps"])
            self._skipped, self._found_inf = int(state.get("skipped", 0)), False


def collect_grads(optimizer) -> List[torch.Tensor]:
    return [p.grad for g in optimizer.param_groups for p in g["params"] if p.grad is not None]
