"""Small host-side utilities: seeding, early stopping, logging, checkpoints, timers."""
from .seeds import same_seeds
from .early_stopping import EarlyStopping
from .checkpoint import save_weights, load_weights, save_resume_state, load_resume_state
from .tb import ScalarLogger
from .timing import CudaStepTimer, nvtx_range

__all__ = ["same_seeds", "EarlyStopping", "save_weights", "load_weights",
           "save_resume_state", "load_resume_state", "ScalarLogger",
           "CudaStepTimer", "nvtx_range"]
