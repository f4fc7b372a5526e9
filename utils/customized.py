"""Compatibility module: ``utils.customized`` of the reference maps onto ``ddpt_b200.utils``."""
from ddpt_b200.utils import same_seeds, EarlyStopping  # noqa: F401
