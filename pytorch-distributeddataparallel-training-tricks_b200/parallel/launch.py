"""Process bootstrap (reference: run.py:22-29).

One process per GPU, ``env://`` rendezvous -- the same contract as the
reference's ``torch.distributed.launch`` line -- plus what it lacked: a gloo path
for CPU-only boxes (BASELINE.json config 1), a single-process default when no
launcher variables are present, a rendezvous timeout, and a clean
``destroy_process_group`` on the way out.
"""
from __future__ import annotations

import datetime
import os
import socket
from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass
class DistEnv:
    rank: int
    local_rank: int
    world_size: int
    device: torch.device
    backend: str

    @property
    def is_main(self) -> bool:
        return self.rank == 0


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def init_distributed(local_rank: int = 0, backend: str = "auto", timeout_s: int = 600) -> DistEnv:
    use_cuda = torch.cuda.is_available()
    if backend == "auto":
        backend = "nccl" if use_cuda else "gloo"
    if backend == "nccl" and not use_cuda:
        raise RuntimeError("backend 'nccl' requested but CUDA is not available; use --backend gloo")
    if use_cuda:
        torch.cuda.set_device(local_rank)
        device = torch.device("cuda", local_rank)
    else:
        device = torch.device("cpu")
    if "RANK" not in os.environ or "WORLD_SIZE" not in os.environ:
        # not under a launcher: become a world of one
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", str(_free_port()))
        os.environ["RANK"], os.environ["WORLD_SIZE"] = "0", "1"
        os.environ.setdefault("LOCAL_RANK", "0")
    if not dist.is_initialized():
        kwargs = dict(backend=backend, init_method="env://",
                      timeout=datetime.timedelta(seconds=timeout_s))
        if backend == "nccl":
            kwargs["device_id"] = device
        dist.init_process_group(**kwargs)
    return DistEnv(rank=dist.get_rank(), local_rank=local_rank, world_size=dist.get_world_size(),
                   device=device, backend=backend)


def shutdown_distributed() -> None:
    if dist.is_available() and dist.is_initialized():
        try:
            dist.barrier()
        except Exception:
            pass
        dist.destroy_process_group()
