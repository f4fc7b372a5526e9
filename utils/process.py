"""Compatibility module: ``utils.process`` of the reference maps onto ``ddpt_b200.engine``."""
from ddpt_b200.engine import iterate_loader  # noqa: F401
