"""Apex-style mixed precision front-end (reference: ``from apex import amp``).

``amp.initialize(model, optimizers=opt, opt_level="O1")`` and
``with amp.scale_loss(loss, opt) as scaled: scaled.backward()`` keep the call
sites of utils/train.py:58 and utils/process.py:26-27 unchanged.
"""
from .frontend import initialize, scale_loss, state_dict, load_state_dict, current_state, reset
from .scaler import LossScaler

__all__ = ["initialize", "scale_loss", "state_dict", "load_state_dict", "current_state",
           "reset", "LossScaler"]
