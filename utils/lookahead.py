"""Compatibility module: ``utils.lookahead`` of the reference maps onto ``ddpt_b200.optim``."""
from ddpt_b200.optim import Lookahead  # noqa: F401
