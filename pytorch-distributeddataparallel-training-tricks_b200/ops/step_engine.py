"""CUDA-graph step engine for the fused Toy_Net path.

``iterate_loader`` (engine/process.py) hands every batch to this engine when the
(model, loss, optimizer) triple is the one it was built for.  One training step is then a single
graph replay containing, in order::

    clear found_inf | zero grad arena | fused forward (+CE metrics) | fused backward writing into
    the gradient arena, with each DDP bucket's fused NVLink all-reduce (x 1/(world*scale), overflow
    flag) forked onto the comm stream as soon as its last gradient is written | join |
    (world == 1: unscale + overflow check) | loss-scale update | Lookahead-SGD (+bf16 weight shadow)

Nothing in the step reads device memory from the host: learning rate, loss scale, overflow flag,
Lookahead counter and the loss / accuracy accumulators all live on the device.  The first call for
a batch size runs eagerly (allocating workspaces), the second captures, later calls replay.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from ..utils.capture import graph_capture
import torch.nn as nn


def _unwrap(model):
    from ..parallel.ddp import DistributedDataParallel
    return (model.module, model) if isinstance(model, DistributedDataParallel) else (model, None)


class StepEngine:
    def __init__(self, model, loss_function, optimizer, use_graph: bool = True):
        from .. import amp
        from ..optim import Lookahead
        self.module, self.ddp = _unwrap(model)
        self.wrapper = model
        self.loss_function = loss_function
        self.optimizer = optimizer
        self.lookahead = optimizer if isinstance(optimizer, Lookahead) else None
        self.sgd = optimizer.optimizer if self.lookahead is not None else optimizer
        self.fused_model = self.module._fused
        self.fused_opt = self.sgd._fused
        self.scaler = amp.current_state().scaler
        self.device = self.fused_opt.device
        self.use_graph = use_graph
        self.metrics = torch.zeros(2, dtype=torch.float32, device=self.device)
        self._graphs: Dict[Tuple[int, bool], torch.cuda.CUDAGraph] = {}
        self._seen: Dict[Tuple[int, bool], int] = {}
        self._static: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = {}
        self._seen_samples = 0
        self.comm_stream = torch.cuda.Stream(device=self.device) if self.ddp is not None else None
        # weight-gradient GEMMs run on a side stream, off the critical path of backward (DDPT_WGRAD_STREAM=0 disables)
        import os
        self.wgrad_stream = (torch.cuda.Stream(device=self.device)
                             if os.environ.get("DDPT_WGRAD_STREAM", "1") != "0" else None)
        # DDPT_MAIN_PRIORITY=1 captures the graph on a high-priority stream (critical path ahead of the side-stream
        # weight gradients).  Measured slower on the same box (1.673 vs 1.638 ms/step): with equal priorities the
        # weight-gradient GEMMs start earlier and fill the SMs the bandwidth-bound passes leave idle.  Off by default.
        self.capture_stream = (torch.cuda.Stream(device=self.device, priority=-1)
                               if os.environ.get("DDPT_MAIN_PRIORITY", "0") == "1" else None)
        self.kernel_launches_per_step: Dict[Tuple[int, bool], int] = {}
        self.graph_replays = 0
        # optional: copy the metric accumulators to pinned host memory after every step (async, no sync)
        self.host_metrics_every_step = False
        self._host_ring = torch.zeros((1024, 2), dtype=torch.float32).pin_memory()
        self._host_slot = 0
        # host batches are staged through two device buffers on a copy stream, so the H2D transfer of step i+1
        # overlaps the compute of step i (the graph's input buffer is then filled by a ~1 us device-to-device copy)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._stage: Dict[Tuple[int, int], Tuple[torch.Tensor, torch.Tensor, Optional[torch.cuda.Event]]] = {}
        self._stage_slot = 0

    # -- eligibility ---------------------------------------------------------------------------------
    def accepts(self, model, loss_function, optimizer, training: bool) -> bool:
        if model is not self.wrapper or loss_function is not self.loss_function:
            return False
        if training and optimizer is not self.optimizer:
            return False
        lf = loss_function
        return (isinstance(lf, nn.CrossEntropyLoss) and lf.reduction == "mean" and lf.weight is None
                and lf.label_smoothing == 0.0 and lf.ignore_index == -100)

    # -- pass bookkeeping (loss / accuracy accumulate on the device) -----------------------------------
    def begin_pass(self) -> None:
        self.metrics.zero_()
        self._seen_samples = 0

    def end_pass(self) -> Tuple[float, float, int]:
        loss_sum, correct = self.metrics.tolist()      # the one host sync of the pass
        return loss_sum, correct, self._seen_samples

    # -- one step ------------------------------------------------------------------------------------------
    def _buckets_for(self):
        return self.ddp.buckets if self.ddp is not None else []

    def _train_body(self, x: torch.Tensor, t: torch.Tensor) -> None:
        fo, fm, scaler = self.fused_opt, self.fused_model, self.scaler
        arena = fo.arena
        batch = x.shape[0]
        scaler.begin_step()
        arena.grad_arena.zero_()
        fm.forward_impl(x, True, head=False)      # loss / accuracy come out of the fused head launch in backward
        grads = fm.arena_grads(arena)
        if self.ddp is not None and self.ddp.world_size > 1:
            backend = self.ddp.reducer.backend
            pending = {b.index: len(b.param_indices) for b in self.ddp.buckets}
            bucket_of = {i: b for b in self.ddp.buckets for i in b.param_indices}
            main = torch.cuda.current_stream()

            def on_ready(pidx: int) -> None:
                b = bucket_of[pidx]
                pending[b.index] -= 1
                if pending[b.index] == 0:
                    if self.wgrad_stream is not None:      # the bucket also holds gradients produced on the side stream
                        ev = torch.cuda.Event()
                        ev.record(self.wgrad_stream)
                        # the stream that will issue the collective must see them: the comm stream for the fused
                        # kernel, the main stream for process-group back ends (they ignore comm_stream)
                        (self.comm_stream if getattr(backend, "uses_comm_stream", False) else main).wait_event(ev)
                    backend.launch_on(self.comm_stream, main, arena, b, scaler)

            fm.backward_impl(batch, grads, target=t, loss_scale=scaler.scale_tensor, on_ready=on_ready,
                             metrics=self.metrics, side_stream=self.wgrad_stream)
            backend.join(self.comm_stream, main)
            if not backend.fuses_unscale:
                fo.unscale(scaler)
        else:
            fm.backward_impl(batch, grads, target=t, loss_scale=scaler.scale_tensor, metrics=self.metrics,
                             side_stream=self.wgrad_stream)
            fo.unscale(scaler)
        scaler.update()
        fo.step(self.sgd, self.lookahead)
        if self.ddp is not None:
            self.ddp._buffers_dirty = True

    def _eval_body(self, x: torch.Tensor, t: torch.Tensor) -> None:
        self.fused_model.forward_impl(x, False, target=t, metrics=self.metrics)

    def run(self, image: torch.Tensor, target: torch.Tensor, training: bool) -> None:
        batch = image.shape[0]
        key = (batch, training)
        if training:
            self.fused_opt.sync_hyper(self.sgd, self.lookahead)
        elif self.ddp is not None and self.ddp._buffers_dirty and self.ddp.world_size > 1:
            self.ddp._sync_buffers()                  # rank 0's BN statistics, once per train->eval switch
            self.ddp._buffers_dirty = False
        static = self._static.get(batch)
        if static is None:
            static = (torch.empty((batch, 1, 28, 28), dtype=image.dtype, device=self.device),
                      torch.empty((batch,), dtype=torch.long, device=self.device))
            self._static[batch] = static
        sx, st = static
        if sx.dtype != image.dtype:
            image = image.to(sx.dtype) if image.dtype != torch.uint8 else image
            if sx.dtype != image.dtype:
                sx = torch.empty_like(sx, dtype=image.dtype)
                self._static[batch] = (sx, st)
                self._graphs.pop(key, None)
                self._seen[key] = 0
        if image.device.type == "cpu" and target.device.type == "cpu":
            slot = self._stage_slot
            self._stage_slot ^= 1
            entry = self._stage.get((batch, slot))
            if entry is None or entry[0].dtype != sx.dtype:
                entry = (torch.empty_like(sx), torch.empty_like(st), None)
            gx, gt, free = entry
            main = torch.cuda.current_stream(self.device)
            if free is not None:
                self.copy_stream.wait_event(free)          # the compute stream has consumed this slot (2 steps ago)
            with torch.cuda.stream(self.copy_stream):
                gx.copy_(image.reshape(sx.shape), non_blocking=True)
                gt.copy_(target, non_blocking=True)
                landed = torch.cuda.Event()
                landed.record(self.copy_stream)
            main.wait_event(landed)
            sx.copy_(gx, non_blocking=True)
            st.copy_(gt, non_blocking=True)
            consumed = torch.cuda.Event()
            consumed.record(main)
            self._stage[(batch, slot)] = (gx, gt, consumed)
        else:
            sx.copy_(image.reshape(sx.shape), non_blocking=True)
            st.copy_(target, non_blocking=True)
        body = self._train_body if training else self._eval_body
        seen = self._seen.get(key, 0)
        graph = self._graphs.get(key)
        if graph is not None:
            graph.replay()
            self.graph_replays += 1
        elif not self.use_graph or seen == 0:
            if training:
                self.fused_opt.prepare_for_capture(self.sgd, self.lookahead)
            body(sx, st)                               # eager: allocates workspaces, a real step
        else:
            if training:
                self.fused_opt.prepare_for_capture(self.sgd, self.lookahead)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with graph_capture(graph, stream=self.capture_stream):
                body(sx, st)
            self._graphs[key] = graph
            graph.replay()
            self.graph_replays += 1
        self._seen[key] = seen + 1
        self._seen_samples += batch
        if self.host_metrics_every_step:
            self._host_ring[self._host_slot % 1024].copy_(self.metrics, non_blocking=True)
            self._host_slot += 1


def attach(model, loss_function, optimizer, use_graph: bool = True) -> bool:
    module, _ = _unwrap(model)
    from ..optim import Lookahead
    inner = optimizer.optimizer if isinstance(optimizer, Lookahead) else optimizer
    if getattr(module, "_fused", None) is None or getattr(inner, "_fused", None) is None:
        return False
    from .. import amp
    scaler = amp.current_state().scaler
    if scaler is None or not scaler.on_device:
        return False
    model._ddpt_step_engine = StepEngine(model, loss_function, optimizer, use_graph)
    return True
