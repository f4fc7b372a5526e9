"""Flat-arena, single-launch Lookahead-SGD (csrc/optim/fused_lookahead_sgd.cu) behind the
``SGD`` / ``Lookahead`` Python API.

``attach(optimizer, models, scaler)`` re-homes the optimizer's parameters onto a fp32 arena (or
adopts the arena a DDP wrapper already created), allocates momentum / slow-weight / bf16 shadow
arenas, and installs a :class:`FusedState` on the inner ``SGD`` so that ``Lookahead.step()`` is
one kernel.  Learning rate, momentum, Lookahead counter and the overflow flag are read from
device memory inside the kernel, so the launch is identical every step and CUDA-graph safe.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import require_native
from ..parallel.arena import FlatArena
from ..parallel.buckets import plan_buckets

HYPER_LR, HYPER_MOMENTUM, HYPER_WD, HYPER_ALPHA, HYPER_DAMP = 0, 1, 2, 3, 4
STATE_STEP, STATE_LA_COUNTER, STATE_SLOW_READY, STATE_K, STATE_TICKET, STATE_SKIPPED = 0, 1, 2, 3, 4, 5


class FusedState:
    def __init__(self, sgd, params: List[torch.nn.Parameter], scaler, arena: Optional[FlatArena] = None,
                 bucket_cap_mb: float = 25.0, shadow_dtype: torch.dtype = torch.bfloat16):
        self.ext = require_native()
        self.params = params
        self.scaler = scaler
        self.device = params[0].device
        self.arena = arena if arena is not None else FlatArena(params, plan_buckets(params, bucket_cap_mb))
        n = self.arena.numel
        self.momentum = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.slow: Optional[torch.Tensor] = None
        self.shadow_dtype = shadow_dtype       # bf16 (O1) or fp16 (O1-fp16): the tensor-core operand format
        self.shadow = torch.empty(n, dtype=shadow_dtype, device=self.device)
        self.hyper = torch.zeros(8, dtype=torch.float32, device=self.device)
        self.state = torch.zeros(8, dtype=torch.int32, device=self.device)
        self._hyper_host = None
        self.zero_grads_in_step = False
        self.refresh_shadow()

    # -- arena adoption (DDP creates its own arena after amp.initialize) ---------------------
    def adopt_arena(self, arena: FlatArena) -> None:
        if arena is self.arena:
            return
        if arena.numel != self.arena.numel:
            if int(self.state[STATE_STEP].item()) != 0:
                raise RuntimeError("cannot adopt an arena with a different layout after training started")
            n = arena.numel
            self.momentum = torch.zeros(n, dtype=torch.float32, device=self.device)
            self.slow = None if self.slow is None else torch.zeros(n, dtype=torch.float32, device=self.device)
            self.shadow = torch.empty(n, dtype=self.shadow_dtype, device=self.device)
        self.arena = arena
        self.refresh_shadow()

    def refresh_shadow(self) -> None:
        self.ext.cast_bf16(self.arena.param_arena, self.shadow)

    def shadow_view(self, pidx: int) -> torch.Tensor:
        """bf16 weights of parameter ``pidx`` in arena (channels-last) physical order, flat."""
        off = self.arena.offset_of[pidx]
        return self.shadow[off: off + self.params[pidx].numel()]

    # -- step --------------------------------------------------------------------------------
    def _push_hyper(self, sgd, lookahead) -> None:
        g = sgd.param_groups[0]
        host = (float(g["lr"]), float(g["momentum"]), float(g["weight_decay"]),
                float(lookahead.alpha) if lookahead is not None else 0.0, float(g["dampening"]),
                int(lookahead.k) if lookahead is not None else 0)
        if host != self._hyper_host:
            self.hyper[:5].copy_(torch.tensor(host[:5], dtype=torch.float32), non_blocking=False)
            self.state[STATE_K:STATE_K + 1].fill_(host[5])
            self._hyper_host = host

    def step(self, sgd, lookahead) -> None:
        if len(sgd.param_groups) != 1:
            raise NotImplementedError("fused Lookahead-SGD supports a single param group")
        if not torch.cuda.is_current_stream_capturing():
            self._push_hyper(sgd, lookahead)
        if lookahead is not None and self.slow is None:
            self.slow = torch.zeros_like(self.momentum)
        found_inf = self.scaler.found_inf_tensor if (self.scaler is not None and self.scaler.on_device) else None
        self.ext.fused_lookahead_sgd(self.arena.param_arena, self.arena.grad_arena, self.momentum,
                                     self.slow if lookahead is not None else None, self.shadow, self.hyper,
                                     self.state, found_inf, bool(sgd.param_groups[0]["nesterov"]),
                                     self.zero_grads_in_step)

    def prepare_for_capture(self, sgd, lookahead) -> None:
        """Materialise lazily created buffers and push hyper-parameters before graph capture."""
        self._push_hyper(sgd, lookahead)
        if lookahead is not None and self.slow is None:
            self.slow = torch.zeros_like(self.momentum)

    def sync_hyper(self, sgd, lookahead) -> None:
        self._push_hyper(sgd, lookahead)

    def zero_grad(self, set_to_none: bool = True) -> None:
        self.arena.zero_grads()

    def unscale(self, scaler, already_unscaled: bool = False) -> None:
        if not already_unscaled:
            self.ext.unscale_arena(self.arena.grad_arena, scaler.state)

    # -- (de)serialisation bridges to the Python optimizer state --------------------------------
    def _views(self, flat: torch.Tensor):
        from ..parallel.arena import _view_like
        return [_view_like(flat, self.arena.offset_of[i], p, self.arena.channels_last)
                for i, p in enumerate(self.params)]

    def sync_to_python(self, sgd, lookahead) -> None:
        st = self.state.tolist()
        if st[STATE_STEP] > 0:
            for p, m in zip(self.params, self._views(self.momentum)):
                sgd.state[p]["momentum_buffer"] = m
        if lookahead is not None:
            for g in sgd.param_groups:
                g["counter"] = st[STATE_LA_COUNTER]
            if st[STATE_SLOW_READY] and self.slow is not None:
                for p, s in zip(self.params, self._views(self.slow)):
                    lookahead.state[p]["slow_param"] = s

    def sync_from_python(self, sgd, lookahead) -> None:
        have_m = all("momentum_buffer" in sgd.state.get(p, {}) and sgd.state[p]["momentum_buffer"] is not None
                     for p in self.params)
        new_state = [0] * 8
        if have_m:
            for p, m in zip(self.params, self._views(self.momentum)):
                src = sgd.state[p]["momentum_buffer"]
                if src.data_ptr() != m.data_ptr():
                    m.copy_(src)
            new_state[STATE_STEP] = 1
        if lookahead is not None:
            if self.slow is None:
                self.slow = torch.zeros_like(self.momentum)
            have_s = all("slow_param" in lookahead.state.get(p, {}) for p in self.params)
            if have_s:
                for p, s in zip(self.params, self._views(self.slow)):
                    src = lookahead.state[p]["slow_param"]
                    if src.data_ptr() != s.data_ptr():
                        s.copy_(src)
                new_state[STATE_SLOW_READY] = 1
            new_state[STATE_LA_COUNTER] = int(sgd.param_groups[0].get("counter", 0))
            new_state[STATE_K] = int(lookahead.k)
        self.state.copy_(torch.tensor(new_state, dtype=torch.int32))
        self._hyper_host = None
        self.refresh_shadow()


def attach(optimizer, models, scaler) -> bool:
    """Install the fused update on ``optimizer`` (an ``SGD`` or a ``Lookahead`` wrapping one)."""
    from ..optim import SGD, Lookahead
    inner = optimizer.optimizer if isinstance(optimizer, Lookahead) else optimizer
    if not isinstance(inner, SGD) or len(inner.param_groups) != 1:
        return False
    params = [p for p in inner.param_groups[0]["params"]]
    if not params or not all(p.is_cuda and p.dtype == torch.float32 for p in params):
        return False
    model_params = [p for m in models for p in m.parameters()]
    if len(model_params) != len(params) or any(a is not b for a, b in zip(model_params, params)):
        return False  # arena layout is defined over the full, ordered parameter list of the model
    from .. import amp
    shadow_dtype = torch.float16 if amp.current_state().compute == "fp16" else torch.bfloat16
    fused = FusedState(inner, params, scaler, shadow_dtype=shadow_dtype)
    inner._fused = fused
    for m in models:
        m._ddpt_fused_optim = fused
        m._ddpt_on_arena = fused.adopt_arena
    return True
