"""Device-side timing and NVTX helpers (the reference has none)."""
from __future__ import annotations

import contextlib
import statistics
from typing import List

import torch


@contextlib.contextmanager
def nvtx_range(name: str, enabled: bool = True):
    """NVTX range that is a no-op without CUDA."""
    on = enabled and torch.cuda.is_available()
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()


class CudaStepTimer:
    """Times regions on the launching stream with CUDA events.

    ``with timer.region():`` records a start/stop event pair; nothing is
    synchronised until :meth:`summary` is called, so timing adds no host stall
    to the loop being measured.
    """

    def __init__(self) -> None:
        self._pairs: List[tuple] = []

    @contextlib.contextmanager
    def region(self):
        if not torch.cuda.is_available():
            yield
            return
        start = torch.cuda.Event(enable_timing=True)
        stop = torch.cuda.Event(enable_timing=True)
        start.record()
        try:
            yield
        finally:
            stop.record()
            self._pairs.append((start, stop))

    def times_ms(self) -> List[float]:
        if self._pairs:
            torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in self._pairs]

    def summary(self) -> dict:
        t = self.times_ms()
        if not t:
            return {"n": 0}
        return {"n": len(t), "mean_ms": statistics.fmean(t), "median_ms": statistics.median(t),
                "min_ms": min(t), "max_ms": max(t)}
