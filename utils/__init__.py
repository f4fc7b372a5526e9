"""Reference-compatible ``utils`` namespace (reference: utils/__init__.py:1-5).

The six public names of the reference, re-exported from the B200-native package
so ``from utils import Toy_Net, Lookahead, ...`` and ``from utils.train import
train`` keep working for code written against the original repository.
"""
from ddpt_b200.utils import same_seeds, EarlyStopping
from ddpt_b200.optim import Lookahead
from ddpt_b200.models import Toy_Net
from ddpt_b200.engine import iterate_loader, train

__all__ = ["same_seeds", "EarlyStopping", "Lookahead", "Toy_Net", "iterate_loader", "train"]
