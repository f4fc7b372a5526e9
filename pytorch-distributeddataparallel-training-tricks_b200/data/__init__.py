"""Input pipeline: MNIST IDX reader, DistributedSampler-parity sharding, batch loaders."""
from .mnist import load_mnist, synthetic_mnist, MNISTArrays
from .sampler import ShardedSampler
from .loader import BatchLoader, make_loaders

__all__ = ["load_mnist", "synthetic_mnist", "MNISTArrays", "ShardedSampler",
           "BatchLoader", "make_loaders"]
