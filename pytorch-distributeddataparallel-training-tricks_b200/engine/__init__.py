"""Training driver and step loop (reference: utils/train.py, utils/process.py)."""
from .process import iterate_loader
from .train import train
from .graphed import GraphedTrainStep

__all__ = ["iterate_loader", "train", "GraphedTrainStep"]
