"""Flat fp32 arenas for parameters and gradients.

Every parameter of the wrapped module becomes a view into one contiguous
"param arena" and its ``.grad`` a view at the same offset of a "grad arena"; the
offsets are the bucket plan of :mod:`.buckets`, so a gradient bucket is simply a
slice of the grad arena (DDP's ``gradient_as_bucket_view`` taken to its
conclusion: no copy-in, no copy-out, one memset for ``zero_grad`` and one kernel
for the optimizer instead of ~150 launches -- SURVEY.md section 7.1).

4-D convolution weights are laid out channels-last *inside* the arena
([C_out][kh][kw][C_in] physically, exposed with the usual [C_out,C_in,kh,kw]
shape via strides).  That is the K-major operand layout the tcgen05 implicit-GEMM
kernels consume directly, and it is what cuDNN prefers for NHWC activations; the
logical values -- and therefore ``state_dict`` -- are unchanged.

When a symmetric-memory allocator is supplied the grad arena is allocated from
NVLink-mapped memory so peers can read/write it in place.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch

from .buckets import Bucket, arena_numel


def _view_like(arena: torch.Tensor, offset: int, ref: torch.Tensor, channels_last: bool) -> torch.Tensor:
    flat = arena[offset: offset + ref.numel()]
    if channels_last and ref.dim() == 4:
        co, ci, kh, kw = ref.shape
        return flat.view(co, kh, kw, ci).permute(0, 3, 1, 2)
    return flat.view(ref.shape)


class FlatArena:
    def __init__(self, params: Sequence[torch.nn.Parameter], buckets: List[Bucket],
                 grad_allocator: Optional[Callable[[int], torch.Tensor]] = None,
                 channels_last_weights: bool = True):
        self.params = list(params)
        self.buckets = buckets
        self.numel = arena_numel(buckets)
        dev = self.params[0].device
        if any(p.dtype != torch.float32 for p in self.params):
            raise TypeError("FlatArena expects fp32 master parameters")
        self.channels_last = channels_last_weights and dev.type == "cuda"
        self.param_arena = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.grad_arena = (grad_allocator(self.numel) if grad_allocator is not None
                           else torch.zeros(self.numel, dtype=torch.float32, device=dev))
        self.grad_arena.zero_()
        self.offset_of = {}
        self.grad_views: List[Optional[torch.Tensor]] = [None] * len(self.params)
        with torch.no_grad():
            for b in buckets:
                for pidx, off in zip(b.param_indices, b.offsets):
                    p = self.params[pidx]
                    view = _view_like(self.param_arena, off, p, self.channels_last)
                    view.copy_(p.data)
                    p.data = view
                    self.grad_views[pidx] = _view_like(self.grad_arena, off, p, self.channels_last)
                    self.offset_of[pidx] = off
        self.bind_grads()

    def bind_grads(self) -> None:
        """(Re)attach ``p.grad`` to the arena views (after ``zero_grad(set_to_none=True)`` etc.)."""
        for p, g in zip(self.params, self.grad_views):
            if p.requires_grad and p.grad is not g:
                p.grad = g

    def zero_grads(self) -> None:
        self.grad_arena.zero_()
        self.bind_grads()

    def bucket_slice(self, b: Bucket) -> torch.Tensor:
        return self.grad_arena[b.start: b.start + b.numel]
