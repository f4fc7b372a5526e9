"""Gradient reducer: counts ready gradients per bucket during backward and launches
the bucket's all-reduce as soon as its last gradient exists.

Two interchangeable back ends implement ``launch(bucket)`` / ``finish()``:

``ProcessGroupBackend``  ``dist.all_reduce(async_op=True)`` on the bucket slice --
    gloo on CPU (plumbing configuration) or NCCL (the measured baseline and the
    correctness oracle for the fused kernel).
``FusedNvlinkBackend``   (parallel/symm.py) our one-shot / two-shot sm_100a kernel
    over NVLink peer or multicast memory on a side stream; the kernel also applies
    1/(world * loss_scale) and raises the found-inf flag, so no separate ``div_`` or
    ``amp_C.multi_tensor_scale`` pass exists on that path.

On CUDA the per-parameter hooks and the bucket state machine run natively
(csrc/reducer/reducer.cpp) on the autograd thread; this Python class is the
portable implementation with the same observable behaviour and the one the CPU
tests exercise.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from .arena import FlatArena
from .buckets import Bucket


class ProcessGroupBackend:
    """All-reduce through ``torch.distributed`` (gloo / nccl)."""

    fuses_unscale = False

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self._pending: List[tuple] = []

    def launch(self, arena: FlatArena, bucket: Bucket, scaler=None) -> None:
        flat = arena.bucket_slice(bucket)
        flat.div_(self.world)
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._pending.append((work, bucket))

    def finish(self) -> None:
        for work, _ in self._pending:
            work.wait()
        self._pending.clear()


class Reducer:
    def __init__(self, arena: FlatArena, backend, find_unused: bool = False):
        self.arena = arena
        self.backend = backend
        self.buckets = arena.buckets
        self._bucket_of = {}
        for b in self.buckets:
            for pidx in b.param_indices:
                self._bucket_of[pidx] = b.index
        self._pending = [0] * len(self.buckets)
        self._next_to_launch = 0
        self._ready = [False] * len(self.buckets)
        self._armed = False
        self._callback_queued = False
        self.require_sync = True           # False inside no_sync()
        self._scaler = None
        self._scaler_consumed = False
        self._hooks = []
        self.stats = {"launches": 0, "steps": 0}
        for pidx, p in enumerate(arena.params):
            if p.requires_grad:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(pidx)))

    # -- amp integration ---------------------------------------------------------------
    def expect_scaled_grads(self, scaler) -> None:
        self._scaler = scaler
        self._scaler_consumed = False

    def consumed_scaler(self) -> bool:
        return self._scaler_consumed

    # -- per-iteration state ----------------------------------------------------------------
    def prepare_for_backward(self) -> None:
        for b in self.buckets:
            self._pending[b.index] = sum(1 for i in b.param_indices if self.arena.params[i].requires_grad)
            self._ready[b.index] = False
        self._next_to_launch = 0
        self._armed = True
        self._callback_queued = False

    def _make_hook(self, pidx: int):
        def hook(param: torch.Tensor) -> None:
            if not (self._armed and self.require_sync):
                return
            view = self.arena.grad_views[pidx]
            if param.grad is not view:
                # autograd produced a fresh tensor (grad was None): fold it into the arena
                view.copy_(param.grad)
                param.grad = view
            if not self._callback_queued:
                self._callback_queued = True
                torch.autograd.Variable._execution_engine.queue_callback(self._finalize)
            bidx = self._bucket_of[pidx]
            self._pending[bidx] -= 1
            if self._pending[bidx] == 0:
                self._ready[bidx] = True
                self._launch_ready_in_order()
        return hook

    def _launch_ready_in_order(self) -> None:
        # buckets go out strictly in index order so every rank issues the same sequence
        while self._next_to_launch < len(self.buckets) and self._ready[self._next_to_launch]:
            b = self.buckets[self._next_to_launch]
            fuse = self.backend.fuses_unscale and self._scaler is not None
            self.backend.launch(self.arena, b, self._scaler if fuse else None)
            if fuse:
                self._scaler_consumed = True
            self.stats["launches"] += 1
            self._next_to_launch += 1

    def _finalize(self) -> None:
        if self._next_to_launch != len(self.buckets):
            missing = [b.index for b in self.buckets if not self._ready[b.index]]
            self._armed = False
            raise RuntimeError(
                f"backward finished but buckets {missing} never became ready: some parameters did "
                "not receive gradients (unused parameters are not supported; freeze them with "
                "requires_grad=False before wrapping)")
        self.backend.finish()
        self._armed = False
        self.stats["steps"] += 1

    def remove_hooks(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks.clear()
