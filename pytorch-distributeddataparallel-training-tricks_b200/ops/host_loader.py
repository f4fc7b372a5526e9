"""Python face of the native pinned-host batch loader (csrc/data/host_loader.cpp).

``PinnedGatherLoader.epoch(indices, num_batches)`` yields ``(images_u8, labels_i64, h2d_bytes)`` where the
tensors are *views of pinned host memory*; the consumer copies them to the device with
``non_blocking=True`` (``iterate_loader`` / the step engine do) and the slot is recycled only after a CUDA
event recorded behind that copy has completed.  Works without a GPU too (plain host buffers), which is
how the CPU tests exercise the threading.
"""
from __future__ import annotations

import os
from typing import Iterator, Tuple

import torch

from . import load_ext


class PinnedGatherLoader:
    def __init__(self, images: torch.Tensor, labels: torch.Tensor, batch_size: int, device=None,
                 depth: int = 6, threads: int = 0):
        ext = load_ext()
        if ext is None or not hasattr(ext, "PinnedGatherLoader"):
            raise RuntimeError("native host loader is not built (python -m ddpt_b200.ops.build)")
        self.images = images.contiguous().view(images.shape[0], -1)
        self.labels = labels.contiguous().to(torch.int64)
        assert self.images.dtype == torch.uint8 and not self.images.is_cuda
        self.batch_size = int(batch_size)
        self.device = torch.device(device) if device is not None else None
        self.cuda = self.device is not None and self.device.type == "cuda" and torch.cuda.is_available()
        row = self.images.shape[1]
        mk = (lambda *s, dt: torch.empty(*s, dtype=dt).pin_memory()) if self.cuda else \
             (lambda *s, dt: torch.empty(*s, dtype=dt))
        self.img_slots = [mk(self.batch_size, row, dt=torch.uint8) for _ in range(depth)]
        self.lbl_slots = [mk(self.batch_size, dt=torch.int64) for _ in range(depth)]
        threads = threads or max(1, min(4, (os.cpu_count() or 2) // 2))
        self._native = ext.PinnedGatherLoader(self.images.data_ptr(), self.labels.data_ptr(), self.images.shape[0],
                                              row, self.batch_size, [t.data_ptr() for t in self.img_slots],
                                              [t.data_ptr() for t in self.lbl_slots], threads)
        self.image_shape = tuple(images.shape[1:])
        self._armed = None          # (indices, num_batches) of a speculatively started epoch
        self.epoch_starts = 0       # native start_epoch calls (own epochs + speculative ones)
        self.adopted_epochs = 0     # epochs that found their gather already running

    def epoch(self, indices: torch.Tensor, num_batches: int, next_epoch=None) -> Iterator[Tuple[torch.Tensor, torch.Tensor, int]]:
        """Yield the batches of one epoch.  ``next_epoch`` (optional callable -> (indices, num_batches) or None): as soon
        as this epoch's last batch has been handed out, the workers start gathering the NEXT epoch's first batches into
        the free slots, so the following ``epoch()`` call finds them ready (no cold start at the epoch boundary -- with 8
        GPUs an MNIST epoch is 8 steps, ~12 ms of device time).  If the next call asks for something else, the
        speculative work is discarded."""
        idx = indices.to(torch.int64).contiguous()
        nb = int(num_batches)
        armed = self._armed
        self._armed = None
        if armed is None or armed[1] != nb or armed[0].shape != idx.shape or not torch.equal(armed[0], idx):
            self._native.start_epoch(idx.data_ptr(), idx.numel(), nb)
            self.epoch_starts += 1
        else:
            self.adopted_epochs += 1          # the speculative gather started at the end of the previous epoch is this epoch
        pending = []            # (slot, event) whose H2D copy may still be in flight
        while True:
            slot, rows = self._native.next()
            if slot < 0:
                break
            img = self.img_slots[slot][:rows].view(rows, 1, *self.image_shape)
            lbl = self.lbl_slots[slot][:rows]
            yield img, lbl, rows * (self.images.shape[1] + 8)
            # the consumer has enqueued its copies on the current stream by the time we are resumed
            if self.cuda:
                ev = torch.cuda.Event()
                ev.record()
                pending.append((slot, ev))
                while pending and (pending[0][1].query() or len(pending) >= len(self.img_slots) - 1):
                    s, e = pending.pop(0)
                    e.synchronize()
                    self._native.release(s)
            else:
                self._native.release(slot)
        for s, e in pending:
            e.synchronize()
            self._native.release(s)
        if next_epoch is not None:
            nxt = next_epoch()
            if nxt is not None:
                nidx = nxt[0].to(torch.int64).contiguous()
                self._native.start_epoch(nidx.data_ptr(), nidx.numel(), int(nxt[1]))      # workers start filling slots now
                self.epoch_starts += 1
                self._armed = (nidx, int(nxt[1]))
