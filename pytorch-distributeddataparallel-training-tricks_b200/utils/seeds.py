"""Reproducibility switch (reference: utils/customized.py:5-13)."""
from __future__ import annotations

import random

import numpy as np
import torch


def same_seeds(seed: int = 18) -> None:
    """Seed every RNG the harness touches and pin library algorithm choice.

    Python ``random``, NumPy, the torch CPU generator and every CUDA generator
    get ``seed``.  cuDNN autotuning is disabled and deterministic algorithms are
    requested, matching the reference; our own sm_100a kernels are
    deterministic by construction except for the split-K weight-gradient
    reduction, which ``DDPT_DETERMINISTIC=1`` switches to an ordered reduce.
    """
    for seeder in (random.seed, np.random.seed, torch.manual_seed):
        seeder(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    cudnn = torch.backends.cudnn
    cudnn.deterministic, cudnn.benchmark = True, False
