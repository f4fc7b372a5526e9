"""Batch loaders (reference: ``DataLoader(..., pin_memory=True, sampler=...)``, utils/train.py:28,30).

Three interchangeable pipelines, all yielding ``(image, target)`` exactly like
the reference's loaders and all honouring ``drop_last=False`` tail batches:

``device``  the whole split lives on the GPU as uint8 (47 MB); a batch is one
            index-gather kernel.  No host work per step at all.
``pinned``  the native C++ loader (csrc/data/host_loader.cpp) gathers rows into
            rotating pinned buffers on worker threads and the copy to the device
            is an async H2D on a side stream, one batch ahead of compute.
``torch``   plain vectorised indexing on the host (CPU path / fallback).

Images are yielded as ``uint8 [B,1,28,28]`` on CUDA (the network's first kernel
does ``/255``) and as ``float32`` in [0,1] on CPU, matching ``ToTensor``.
"""
from __future__ import annotations

import math
from typing import Iterator, Optional, Tuple

import torch

from .mnist import MNISTArrays
from .sampler import ShardedSampler


class BatchLoader:
    def __init__(self, data: MNISTArrays, batch_size: int, sampler: Optional[ShardedSampler] = None,
                 device: Optional[torch.device] = None, mode: str = "auto",
                 as_float: Optional[bool] = None, max_batches: Optional[int] = None):
        self.data = data
        self.batch_size = int(batch_size)
        self.sampler = sampler
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        if mode == "auto":
            mode = "device" if self.device.type == "cuda" else "torch"
        if mode in ("device", "pinned") and self.device.type != "cuda":
            raise ValueError(f"loader mode {mode!r} needs a CUDA device")
        self.mode = mode
        self.as_float = (self.device.type != "cuda") if as_float is None else as_float
        self.max_batches = max_batches
        self._dev_images = self._dev_labels = None
        self._native = None
        if mode == "device":
            self._dev_images = data.images.to(self.device).view(-1, 1, 28, 28)
            self._dev_labels = data.labels.to(self.device)
        elif mode == "pinned":
            from ..ops import host_loader
            self._native = host_loader.PinnedGatherLoader(data.images, data.labels, self.batch_size,
                                                          self.device)
        # host-visible counters (bench.py reports them)
        self.h2d_bytes = 0
        self.samples_yielded = 0
        self.quiet = False          # True: iterate_loader shows no progress bar
        self.prefetch_next_epoch = True   # pinned mode: start gathering epoch e+1 while epoch e's last step runs

    # -- sizing -----------------------------------------------------------
    def _num_samples(self) -> int:
        return len(self.sampler) if self.sampler is not None else len(self.data)

    def __len__(self) -> int:
        n = math.ceil(self._num_samples() / self.batch_size)
        return n if self.max_batches is None else min(n, self.max_batches)

    def _epoch_indices(self) -> torch.Tensor:
        return self.sampler.indices() if self.sampler is not None else torch.arange(len(self.data))

    # -- iteration ----------------------------------------------------------
    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        idx = self._epoch_indices()
        nb = len(self)
        if self.mode == "device":
            idx_dev = idx.to(self.device, non_blocking=True)
            self.h2d_bytes += idx.numel() * 8
            for b in range(nb):
                sel = idx_dev[b * self.batch_size:(b + 1) * self.batch_size]
                img = self._dev_images.index_select(0, sel)
                self.samples_yielded += int(sel.numel())
                yield (img.float().div_(255.0) if self.as_float else img), self._dev_labels.index_select(0, sel)
        elif self.mode == "pinned":
            def next_epoch():
                # speculation: the caller will ask for the following epoch with the same batch budget
                if self.sampler is None or not self.prefetch_next_epoch:
                    return None
                return self.sampler.indices_for(self.sampler.epoch + 1), nb
            for img, tgt, nbytes in self._native.epoch(idx, nb, next_epoch):
                self.h2d_bytes += nbytes
                self.samples_yielded += int(tgt.numel())
                yield (img.float().div_(255.0) if self.as_float else img), tgt
        else:
            images, labels = self.data.images, self.data.labels
            for b in range(nb):
                sel = idx[b * self.batch_size:(b + 1) * self.batch_size]
                img = images.index_select(0, sel).view(-1, 1, 28, 28)
                tgt = labels.index_select(0, sel)
                self.samples_yielded += int(sel.numel())
                if self.device.type == "cuda":
                    img = img.pin_memory().to(self.device, non_blocking=True)
                    tgt = tgt.pin_memory().to(self.device, non_blocking=True)
                    self.h2d_bytes += img.numel() + tgt.numel() * 8
                yield (img.float().div_(255.0) if self.as_float else img), tgt


def make_loaders(args, device: torch.device, world_size: int, rank: int):
    """Train/valid loaders with the reference's semantics (train sharded, valid redundant)."""
    from .mnist import load_mnist, synthetic_mnist
    if getattr(args, "synthetic", False):
        train, valid = synthetic_mnist(60000, seed=1), synthetic_mnist(10000, seed=2)
    else:
        train, valid = load_mnist(args.data_path, True), load_mnist(args.data_path, False)
    mode = getattr(args, "loader", "auto")
    max_batches = getattr(args, "max_steps", None)
    train_sampler = ShardedSampler(len(train), num_replicas=world_size, rank=rank)
    valid_sampler = None
    if getattr(args, "shard_valid", False) and world_size > 1:
        valid_sampler = ShardedSampler(len(valid), num_replicas=world_size, rank=rank, shuffle=False)
    train_loader = BatchLoader(train, args.batch_size, train_sampler, device, mode, max_batches=max_batches)
    valid_loader = BatchLoader(valid, args.batch_size, valid_sampler, device, mode, max_batches=max_batches)
    return train_loader, valid_loader, train_sampler
