"""amp_C-equivalent multi-tensor kernels (csrc/amp/multi_tensor.cu)."""
from __future__ import annotations

from typing import List, Optional

import torch

from . import require_native
from ..amp.scaler import LossScaler


def multi_tensor_scale(src: List[torch.Tensor], dst: List[torch.Tensor], scale: torch.Tensor,
                       found_inf: Optional[torch.Tensor] = None) -> None:
    require_native().multi_tensor_scale(src, dst, scale, found_inf)


def multi_tensor_axpby(x: List[torch.Tensor], y: List[torch.Tensor], out: List[torch.Tensor], a: float, b: float,
                       arg_to_check: int = -1, found_inf: Optional[torch.Tensor] = None) -> None:
    require_native().multi_tensor_axpby(x, y, out, a, b, arg_to_check, found_inf)


def unscale_and_check(grads: List[torch.Tensor], scaler_state: torch.Tensor) -> None:
    """grads *= 1/scale in place; raise the found-inf word of the scaler state on overflow."""
    require_native().multi_tensor_scale(grads, grads, scaler_state[1:2], scaler_state[2:3])


def unscale_arena(grad_arena: torch.Tensor, scaler_state: torch.Tensor) -> None:
    require_native().unscale_arena(grad_arena, scaler_state)


def update_scale(scaler_state: torch.Tensor, growth_interval: int = LossScaler.GROWTH_INTERVAL,
                 min_scale: float = LossScaler.MIN_SCALE, max_scale: float = LossScaler.MAX_SCALE) -> None:
    require_native().update_scale(scaler_state, growth_interval, min_scale, max_scale)
