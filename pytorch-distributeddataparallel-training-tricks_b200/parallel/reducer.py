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

    # stream-explicit protocol used by the CUDA-graph step engine: the collective is issued synchronously on the
    # current (capturing) stream -- no overlap, but capturable and correct; this is the NCCL baseline path
    uses_comm_stream = False             # the step engine makes the *main* stream wait for side-stream gradients

    def launch_on(self, comm_stream, main_stream, arena: FlatArena, bucket: Bucket, scaler=None) -> None:
        flat = arena.bucket_slice(bucket)
        flat.div_(self.world)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)

    def join(self, comm_stream, main_stream) -> None:
        return None


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


class NativeReducerAdapter:
    """The C++ reducer (csrc/reducer/reducer.cpp) behind the same interface as :class:`Reducer`."""

    def __init__(self, ext, arena: FlatArena, backend, process_group=None):
        self.arena = arena
        self.backend = backend
        self.buckets = arena.buckets
        self._native = ext.NativeReducer(list(arena.params), list(arena.grad_views), arena.grad_arena,
                                         [list(b.param_indices) for b in self.buckets],
                                         [b.start for b in self.buckets], [b.numel for b in self.buckets])
        if getattr(backend, "fuses_unscale", False):
            r = backend.region
            self._native.use_fused(r.peer_ptrs, r.signal_ptrs, r.mc_ptr, r.scratch_ptrs, r.mc_scratch_ptr, r.rank,
                                   backend.algo, backend.wire, backend.blocks, backend._unit)
        else:
            pg = process_group if process_group is not None else dist.distributed_c10d._get_default_group()
            self._native.use_process_group(pg)
        self.native = True

    def expect_scaled_grads(self, scaler) -> None:
        fuse = getattr(self.backend, "fuses_unscale", False) and scaler is not None and scaler.on_device
        self._native.expect_scaled_grads(scaler.state if fuse else None)

    def consumed_scaler(self) -> bool:
        return self._native.consumed_scaler()

    def prepare_for_backward(self) -> None:
        self._native.prepare_for_backward()

    @property
    def require_sync(self) -> bool:
        return self._native.require_sync

    @require_sync.setter
    def require_sync(self, value: bool) -> None:
        self._native.require_sync = bool(value)

    @property
    def stats(self) -> dict:
        return {"launches": self._native.launches, "steps": self._native.steps,
                "folded_copies": self._native.folded_copies}

    def remove_hooks(self) -> None:
        self._native.require_sync = False


def make_reducer(arena: FlatArena, backend, process_group=None, prefer_native: bool = True):
    """Native C++ reducer when the extension is importable (CPU or CUDA), else the Python one."""
    import os
    if prefer_native and os.environ.get("DDPT_PY_REDUCER") != "1":
        from ..ops import load_ext
        ext = load_ext()
        if ext is not None and hasattr(ext, "NativeReducer"):
            return NativeReducerAdapter(ext, arena, backend, process_group)
    return Reducer(arena, backend)
