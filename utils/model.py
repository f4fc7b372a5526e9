"""Compatibility module: ``utils.model`` of the reference maps onto ``ddpt_b200.models``."""
from ddpt_b200.models import Toy_Net  # noqa: F401
