"""Python face of the sm_100a extension (csrc/ -> _ext/_ddpt_C*.so).

The extension is built in-tree by ``__graft_entry__.build()`` / ``python -m
ddpt_b200.ops.build`` (nvcc ``-gencode arch=compute_100a,code=sm_100a``) and
loaded lazily here.  On a machine with a CUDA device the extension is mandatory:
if it cannot be loaded, :func:`require_native` raises instead of silently falling
back to library kernels.  Without CUDA (the CPU plumbing configuration) every
``attach_*`` is a no-op and the pure-PyTorch paths run.
"""
from __future__ import annotations

import importlib
import os
from typing import Optional

import torch

_EXT = None
_EXT_ERROR: Optional[BaseException] = None
_EXT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), os.pardir, "_ext")


def ext_dir() -> str:
    return os.path.abspath(_EXT_DIR)


def load_ext(build_if_missing: bool = False):
    """Import ``_ddpt_C`` from the in-tree ``_ext`` directory (optionally building it)."""
    global _EXT, _EXT_ERROR
    if _EXT is not None:
        return _EXT
    import sys
    d = ext_dir()
    if d not in sys.path:
        sys.path.insert(0, d)
    try:
        _EXT = importlib.import_module("_ddpt_C")
        _EXT_ERROR = None
    except BaseException as exc:  # ImportError, OSError (missing libs), ...
        _EXT_ERROR = exc
        if build_if_missing:
            from . import build as _build
            _build.build_all(verbose=False)
            importlib.invalidate_caches()
            _EXT = importlib.import_module("_ddpt_C")
            _EXT_ERROR = None
    return _EXT


def native_available() -> bool:
    """True when CUDA is present and the sm_100a extension imported."""
    if not torch.cuda.is_available():
        return False
    if os.environ.get("DDPT_DISABLE_NATIVE") == "1":
        return False
    return load_ext() is not None


def require_native():
    ext = load_ext()
    if ext is None:
        raise RuntimeError(
            "ddpt_b200: the sm_100a extension (_ext/_ddpt_C*.so) is not loadable "
            f"({_EXT_ERROR!r}); build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    return ext


def deterministic() -> bool:
    """``DDPT_DETERMINISTIC=1`` (or ``same_seeds(..., deterministic=True)``): every reduction of the fused step runs
    in a fixed order -- partial rows + finalize kernels / slabs + reduction kernels instead of floating-point atomics
    -- so two runs of the same program produce bit-identical parameters (reference: utils/customized.py:12-13,
    ``cudnn.deterministic = True``).  Costs a few percent of step time."""
    return os.environ.get("DDPT_DETERMINISTIC", "0") == "1"


def attach_fused_model(model: torch.nn.Module, compute: Optional[str]) -> bool:
    """Give ``model`` its fused sm_100a forward/backward if one exists for its type."""
    from ..models import Toy_Net
    if isinstance(model, Toy_Net) and compute in ("bf16", "fp16", "fp8"):
        from . import toynet_fused
        return toynet_fused.attach(model, compute)
    return False


def attach_fused_optimizer(optimizer, models, scaler) -> bool:
    from . import fused_optim
    return fused_optim.attach(optimizer, models, scaler)


def attach_step_engine(model, loss_function, optimizer, use_graph: bool = True) -> bool:
    if not native_available():
        return False
    from . import step_engine
    return step_engine.attach(model, loss_function, optimizer, use_graph=use_graph)
