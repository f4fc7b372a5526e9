"""A whole training step of an ARBITRARY model as one CUDA graph.

The Toy_Net step engine (ops/step_engine.py) captures a hand-fused step; this is the general form for any
``nn.Module`` that trains through autograd:

    optimizer.zero_grad()  ->  model(x)  ->  loss  ->  amp.scale_loss(...).backward()  ->  optimizer.step()

captured once and replayed per batch.  Everything the framework adds to such a step is capture-safe by construction --
the DDP reducer launches its fused NVLink all-reduce kernels from autograd hooks onto a side stream that is forked
from and joined back into the capturing stream (csrc/reducer/reducer.cpp); barrier epochs, loss scale, overflow flag,
learning rate and Lookahead counter live in device memory (no host reads); the buffer broadcast is our kernel -- so the
replay contains forward, backward, bucketed communication and the optimizer with no Python and no launch gaps.  On a
B200 a ResNet-50 step at 64 images per GPU is launch-bound in eager mode; the graph is what makes the per-GPU
batch of BASELINE.json config 4 run at hardware speed.

Stream rule (the same one torch documents for its own DDP under graph capture): autograd runs a parameter's
AccumulateGrad node on the stream that was current when the node was created, and the DDP reducer keeps those nodes
alive (its hooks hang off them).  Construct the DDP wrapper under the stream the step will be captured on::

    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        net = DistributedDataParallel(model, device_ids=[rank], comm="fused")
    step = GraphedTrainStep(net, loss_fn, optimizer, x, y, stream=s)

Reference call sites replaced: the body of the batch loop in utils/process.py:17-28."""
from __future__ import annotations

from typing import Callable, Optional

import torch
from ..utils.capture import graph_capture

from .. import amp


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, loss_fn: Callable, optimizer, example_input: torch.Tensor,
                 example_target: torch.Tensor, warmup: int = 3, autocast_dtype: Optional[torch.dtype] = None,
                 use_graph: bool = True, stream: Optional[torch.cuda.Stream] = None):
        self.model, self.loss_fn, self.optimizer = model, loss_fn, optimizer
        self.autocast_dtype = autocast_dtype
        self.static_x = example_input.clone()
        self.static_y = example_target.clone()
        self.loss = torch.zeros((), device=example_input.device)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.replays = 0
        from ..optim import Lookahead
        self._lookahead = optimizer if isinstance(optimizer, Lookahead) else None
        self._sgd = optimizer.optimizer if self._lookahead is not None else optimizer
        self._fused = getattr(self._sgd, "_fused", None)
        # eager warm-up on a side stream: allocator pools, cuDNN plans, momentum / slow-weight buffers, tile plans
        side = stream if stream is not None else torch.cuda.Stream(device=example_input.device)
        self.stream = side
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(example_input.device)
        if use_graph:
            if self._fused is not None:
                self._fused.prepare_for_capture(self._sgd, self._lookahead)
            self.graph = torch.cuda.CUDAGraph()
            with graph_capture(self.graph, stream=side):
                self._body()

    def _body(self) -> None:
        self.optimizer.zero_grad()
        if self.autocast_dtype is not None:
            with torch.autocast("cuda", dtype=self.autocast_dtype):
                out = self.model(self.static_x)
        else:
            out = self.model(self.static_x)
        loss = self.loss_fn(out.float(), self.static_y)
        with amp.scale_loss(loss, self.optimizer) as scaled:
            scaled.backward()
        self.optimizer.step()
        self.loss.copy_(loss.detach())

    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """Run one optimizer step on (x, y); returns the (device-resident) loss of this step."""
        self.static_x.copy_(x, non_blocking=True)
        self.static_y.copy_(y, non_blocking=True)
        if self.graph is None:
            self._body()
        else:
            if self._fused is not None:
                self._fused.sync_hyper(self._sgd, self._lookahead)      # scheduler changes reach the device-side lr
            self.graph.replay()
            self.replays += 1
        return self.loss
