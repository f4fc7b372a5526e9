"""Rank sharding with the exact index stream of ``torch.utils.data.DistributedSampler``.

Reference call site: utils/train.py:26 (``DistributedSampler(train_set)`` with
defaults shuffle=True, seed=0, drop_last=False) and ``set_epoch`` at :70.  The
permutation is ``randperm(N, generator=seed+epoch)`` on the CPU generator, padded
by wrap-around to a multiple of the world size, then strided ``[rank::world]``.
Keeping the stream bit-identical is what lets the epoch-0 oracle in SURVEY.md
section 4 be reproduced.
"""
from __future__ import annotations

import math
from typing import Iterator, Optional

import torch


class ShardedSampler:
    def __init__(self, num_samples: int, num_replicas: Optional[int] = None,
                 rank: Optional[int] = None, shuffle: bool = True, seed: int = 0,
                 drop_last: bool = False):
        if num_replicas is None or rank is None:
            import torch.distributed as dist
            live = dist.is_available() and dist.is_initialized()
            num_replicas = (dist.get_world_size() if live else 1) if num_replicas is None else num_replicas
            rank = (dist.get_rank() if live else 0) if rank is None else rank
        if not 0 <= rank < num_replicas:
            raise ValueError(f"rank {rank} outside [0, {num_replicas})")
        self.dataset_len = int(num_samples)
        self.num_replicas, self.rank = int(num_replicas), int(rank)
        self.shuffle, self.seed, self.drop_last = shuffle, seed, drop_last
        self.epoch = 0
        self._cache = {}            # epoch -> indices; filled one epoch ahead by a background thread
        self._prefetch = None
        import threading
        self._lock = threading.Lock()
        if drop_last and self.dataset_len % self.num_replicas:
            self.num_samples = math.ceil((self.dataset_len - self.num_replicas) / self.num_replicas)
        else:
            self.num_samples = math.ceil(self.dataset_len / self.num_replicas)
        self.total_size = self.num_samples * self.num_replicas

    def set_epoch(self, epoch: int) -> None:
        self.epoch = int(epoch)

    def indices(self) -> torch.Tensor:
        """This rank's sample indices for the current epoch (int64, CPU).  The next epoch's permutation is
        computed on a background thread while this epoch trains (``randperm`` releases the GIL), so the host
        never stalls the device at an epoch boundary."""
        return self.indices_for(self.epoch)

    def indices_for(self, epoch: int) -> torch.Tensor:
        """Indices of ``epoch`` (cached); also starts computing ``epoch + 1`` in the background."""
        import threading
        with self._lock:
            pending = self._prefetch
        if pending is not None:
            pending.join()
        with self._lock:
            self._prefetch = None
            cur = self._cache.get(epoch)
        if cur is None:
            cur = self._indices_for(epoch)
        with self._lock:
            self._cache = {e: v for e, v in self._cache.items() if e >= epoch}     # drop the past
            self._cache[epoch] = cur
            nxt = epoch + 1
            if self.shuffle and nxt not in self._cache:
                def work():
                    val = self._indices_for(nxt)
                    with self._lock:
                        self._cache[nxt] = val
                self._prefetch = threading.Thread(target=work, daemon=True)
                self._prefetch.start()
        return cur

    def _indices_for(self, epoch: int) -> torch.Tensor:
        if self.shuffle:
            g = torch.Generator()
            g.manual_seed(self.seed + epoch)
            order = torch.randperm(self.dataset_len, generator=g)
        else:
            order = torch.arange(self.dataset_len)
        if self.drop_last:
            order = order[: self.total_size]
        else:
            pad = self.total_size - order.numel()
            if pad > 0:
                reps = math.ceil(pad / max(order.numel(), 1))
                order = torch.cat([order, order.repeat(reps)[:pad]])
        return order[self.rank:self.total_size:self.num_replicas].contiguous()

    def __iter__(self) -> Iterator[int]:
        return iter(self.indices().tolist())

    def __len__(self) -> int:
        return self.num_samples
