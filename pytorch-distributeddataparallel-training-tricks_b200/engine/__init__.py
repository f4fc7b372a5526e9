"""Training driver and step loop (reference: utils/train.py, utils/process.py)."""
from .process import iterate_loader
from .train import train

__all__ = ["iterate_loader", "train"]
