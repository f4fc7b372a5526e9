"""Optimizers and LR schedules: Nesterov SGD, Lookahead, warm-up lambda, plateau decay."""
from .sgd import SGD
from .lookahead import Lookahead
from .schedulers import LambdaLR, ReduceLROnPlateau, warmup_lambda

__all__ = ["SGD", "Lookahead", "LambdaLR", "ReduceLROnPlateau", "warmup_lambda"]
