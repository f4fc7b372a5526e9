"""Compatibility module: ``utils.train`` of the reference maps onto ``ddpt_b200.engine``."""
from ddpt_b200.engine import train  # noqa: F401
