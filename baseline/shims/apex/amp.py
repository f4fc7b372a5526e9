import contextlib
import functools
import os

import torch

_DTYPE = {"fp16": torch.float16, "bf16": torch.bfloat16}[os.environ.get("REF_AMP_DTYPE", "fp16")]


class _State:
    scale = 2.0 ** 16
    clean = 0
    enabled = False


def initialize(models, optimizers=None, opt_level="O1", **_):
    _State.enabled = opt_level != "O0"
    model_list = models if isinstance(models, (list, tuple)) else [models]
    for m in model_list:
        inner = m.forward

        @functools.wraps(inner)
        def fwd(*a, __inner=inner, **k):
            with torch.autocast("cuda", dtype=_DTYPE, enabled=_State.enabled):
                return __inner(*a, **k)

        m.forward = fwd
    return (models, optimizers) if optimizers is not None else models


@contextlib.contextmanager
def scale_loss(loss, optimizer, **_):
    if not _State.enabled:
        yield loss
        return
    yield loss.float() * _State.scale
    params = [p for g in optimizer.param_groups for p in g["params"] if p.grad is not None]
    grads = [p.grad for p in params]
    found = torch.zeros(1, device=grads[0].device)
    inv = torch.full((1,), 1.0 / _State.scale, device=grads[0].device)
    torch._amp_foreach_non_finite_check_and_unscale_(grads, found, inv)   # == amp_C.multi_tensor_scale
    if found.item():                                                        # Apex reads the flag on the host too
        _State.scale = max(_State.scale / 2.0, 1.0)
        _State.clean = 0
        real = optimizer.step

        def skip(*a, **k):
            optimizer.step = real

        optimizer.step = skip
    else:
        _State.clean += 1
        if _State.clean >= 2000:
            _State.scale = min(_State.scale * 2.0, 2.0 ** 24)
            _State.clean = 0
