import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# single-GPU multi-"rank" simulations launch co-operating kernels on many streams: give every stream its own
# hardware queue so that one spinning kernel cannot be queued in front of the kernel it waits for.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")
    config.addinivalue_line("markers", "slow: takes more than ~30 s on CPU")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def reference_root():
    if not os.path.isdir(os.path.join(REFERENCE, "utils")):
        pytest.skip("reference checkout not present")
    return REFERENCE


def load_reference_module(name):
    """Import /root/reference/utils/<name>.py without triggering utils/__init__ (needs apex)."""
    import importlib.util
    path = os.path.join(REFERENCE, "utils", f"{name}.py")
    spec = importlib.util.spec_from_file_location(f"_ref_{name}", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
