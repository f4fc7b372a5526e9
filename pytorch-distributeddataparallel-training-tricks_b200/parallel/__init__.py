"""Data parallelism: bucket planning, gradient reducer, DDP wrapper, NVLink symmetric-memory comm."""
from .buckets import plan_buckets, Bucket
from .ddp import DistributedDataParallel
from .launch import init_distributed, shutdown_distributed, DistEnv

__all__ = ["plan_buckets", "Bucket", "DistributedDataParallel", "init_distributed",
           "shutdown_distributed", "DistEnv"]
