"""``initialize`` / ``scale_loss`` with Apex O1 semantics on a Blackwell compute path.

What Apex O1 did on the reference stack: keep fp32 master
weights, cast inputs/weights of conv/linear to fp16 per call, compute losses in
fp32, multiply the loss by a dynamic scale, and after backward unscale + inf/nan
check every gradient with ``amp_C`` multi-tensor kernels, skipping the step on
overflow.

What happens here for the same two calls:

``initialize``  picks the compute dtype (bf16 by default on B200; fp16 and
                block-scaled fp8 selectable), attaches the fused sm_100a forward
                /backward to models that have one (Toy_Net), wraps other models'
                ``forward`` in autocast, moves the optimizer onto flat arenas
                with the single-launch update, and creates the device-resident
                :class:`LossScaler`.
``scale_loss``  yields ``loss.float() * scale``; on exit the gradients are
                unscaled and checked -- by the fused all-reduce kernel itself
                when the model is wrapped in our DDP (no extra pass), otherwise
                by one multi-tensor kernel -- and the scale is updated on device.
                An overflow makes the next ``optimizer.step()`` a no-op,
                including the Lookahead counter.
"""
from __future__ import annotations

import contextlib
import functools
from dataclasses import dataclass, field
from typing import Any, List, Optional, Union

import torch

from .scaler import LossScaler, collect_grads

_LEVELS = {
    "O0": dict(compute=None, dynamic=False),
    "O1": dict(compute="bf16", dynamic=True),
    "O1-fp16": dict(compute="fp16", dynamic=True),
    "O1-fp8": dict(compute="fp8", dynamic=True),
}


@dataclass
class AmpState:
    opt_level: str = "O0"
    compute: Optional[str] = None          # None | "bf16" | "fp16" | "fp8"
    scaler: Optional[LossScaler] = None
    optimizers: List[Any] = field(default_factory=list)
    models: List[torch.nn.Module] = field(default_factory=list)
    _reducers: List[Any] = field(default_factory=list)  # (weakref(reducer), weakref(module)) of live DDP wrappers
    enabled: bool = False

    def register_reducer(self, reducer, module=None) -> None:
        """A DDP wrapper announces its reducer.  Re-wrapping the same module replaces the old entry and dead
        wrappers drop out, so a stale reducer can never veto ``fused_by_reducer`` (which would unscale twice)."""
        import weakref
        mref = weakref.ref(module) if module is not None else None
        keep = []
        for rref, mr in self._reducers:
            if rref() is None or (mref is not None and mr is not None and mr() is module):
                continue
            keep.append((rref, mr))
        keep.append((weakref.ref(reducer), mref))
        self._reducers = keep

    @property
    def reducers(self) -> List[Any]:
        live = [(r, m) for r, m in self._reducers if r() is not None]
        self._reducers = live
        return [r() for r, _ in live]


_STATE = AmpState()


def current_state() -> AmpState:
    return _STATE


def reset() -> None:
    """Forget everything ``initialize`` did (tests / repeated runs in one process)."""
    global _STATE
    _STATE = AmpState()


def _torch_dtype(compute: Optional[str]):
    return {None: None, "bf16": torch.bfloat16, "fp16": torch.float16, "fp8": torch.bfloat16}[compute]


def _wrap_forward_autocast(model: torch.nn.Module, dtype: torch.dtype) -> None:
    inner = model.forward

    @functools.wraps(inner)
    def autocast_forward(*args, **kwargs):
        with torch.autocast(device_type="cuda", dtype=dtype):
            return inner(*args, **kwargs)

    model.forward = autocast_forward  # instance attribute; state_dict is untouched
    model._ddpt_autocast = True


def initialize(models, optimizers=None, opt_level: str = "O1", loss_scale: Union[str, float] = "dynamic",
               verbosity: int = 0, fused: bool = True, **_ignored):
    """Prepare model(s) and optimizer(s) for mixed precision; returns them in the shapes given."""
    if opt_level not in _LEVELS:
        raise ValueError(f"opt_level must be one of {sorted(_LEVELS)}, got {opt_level!r}")
    model_list = list(models) if isinstance(models, (list, tuple)) else [models]
    opt_list = [] if optimizers is None else (list(optimizers) if isinstance(optimizers, (list, tuple)) else [optimizers])

    params = [p for m in model_list for p in m.parameters()]
    device = params[0].device if params else torch.device("cpu")
    level = dict(_LEVELS[opt_level])
    if device.type != "cuda" and level["compute"] is not None:
        # CPU/gloo plumbing configuration (BASELINE.json config 1): fp32, unit static scale
        level = dict(compute=None, dynamic=False)
        opt_level = "O0"

    dynamic = level["dynamic"] and loss_scale == "dynamic"
    init_scale = None if loss_scale == "dynamic" else float(loss_scale)
    st = _STATE
    st.opt_level, st.compute, st.enabled = opt_level, level["compute"], True
    st.scaler = LossScaler(device, dynamic=dynamic, init_scale=init_scale)
    st.models, st.optimizers = model_list, opt_list

    if device.type == "cuda":
        from .. import ops
        for m in model_list:
            attached = False
            if fused and ops.native_available():
                attached = ops.attach_fused_model(m, st.compute)
            if not attached and st.compute is not None and not getattr(m, "_ddpt_autocast", False):
                _wrap_forward_autocast(m, _torch_dtype(st.compute))
        if fused and ops.native_available():
            for opt in opt_list:
                ops.attach_fused_optimizer(opt, model_list, st.scaler)
    if verbosity:
        print(f"amp: opt_level={opt_level} compute={st.compute} loss_scale="
              f"{'dynamic' if dynamic else st.scaler.loss_scale()}")

    out_models = models if isinstance(models, (list, tuple)) else model_list[0]
    if optimizers is None:
        return out_models
    out_opts = optimizers if isinstance(optimizers, (list, tuple)) else opt_list[0]
    return out_models, out_opts


def _arm_skip(optimizer, scaler: LossScaler) -> None:
    """Make the next ``optimizer.step()`` a no-op, once (host-side path only)."""
    real_step = optimizer.step

    def skip_step(*args, **kwargs):
        optimizer.step = real_step
        return None

    optimizer.step = skip_step


@contextlib.contextmanager
def scale_loss(loss: torch.Tensor, optimizers, loss_id: int = 0, delay_unscale: bool = False):
    st = _STATE
    opt_list = list(optimizers) if isinstance(optimizers, (list, tuple)) else [optimizers]
    if not st.enabled or st.scaler is None:
        yield loss
        return
    scaler = st.scaler
    scaler.begin_step()
    reducers = st.reducers
    for r in reducers:
        r.expect_scaled_grads(scaler)
    yield scaler.scale(loss)
    if delay_unscale:
        return
    fused_by_reducer = bool(reducers) and all(r.consumed_scaler() for r in reducers)
    for opt in opt_list:
        fused_opt = getattr(getattr(opt, "optimizer", opt), "_fused", None)
        if fused_opt is not None:
            fused_opt.unscale(scaler, already_unscaled=fused_by_reducer)
        else:
            scaler.unscale_(collect_grads(opt), already_unscaled=fused_by_reducer)
            if scaler.dynamic and scaler.found_inf_host():
                _arm_skip(opt, scaler)
    scaler.update()


def state_dict() -> dict:
    return {"opt_level": _STATE.opt_level,
            "scaler": _STATE.scaler.state_dict() if _STATE.scaler is not None else None}


def load_state_dict(state: dict) -> None:
    if _STATE.scaler is not None and state.get("scaler") is not None:
        _STATE.scaler.load_state_dict(state["scaler"])
