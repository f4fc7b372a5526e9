"""B200-native data-parallel training harness.

Same capabilities and entrypoints as
Lance0218/Pytorch-DistributedDataParallel-Training-Tricks (``run.py``,
``utils.{same_seeds, EarlyStopping, Lookahead, Toy_Net, iterate_loader, train}``)
re-designed for Blackwell: flat parameter/gradient arenas, a fused
all-reduce+unscale+overflow kernel over NVLink peer / multicast memory, a
single-launch Lookahead-SGD update, tcgen05/TMEM + TMA GEMM/conv kernels and a
CUDA-graph captured training step with no per-step host synchronisation.

Import name: ``ddpt_b200`` (see ``ddpt_b200/__init__.py`` at the repo root).
"""
__version__ = "0.1.0"

from . import config  # noqa: F401  (cheap, no torch import side effects beyond argparse)
