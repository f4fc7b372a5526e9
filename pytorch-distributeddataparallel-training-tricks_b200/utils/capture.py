"""CUDA-graph capture with the host-side hazards fenced off."""
import contextlib
import gc

import torch


@contextlib.contextmanager
def graph_capture(graph: "torch.cuda.CUDAGraph", stream=None):
    """``torch.cuda.graph`` for a training step, robust against unrelated host activity.

    A capture in the default ("global") error mode is invalidated by ANY potentially unsafe CUDA call made while it is
    open -- including a ``cudaFree`` / ``cudaGraphDestroy`` issued by Python's cyclic garbage collector when it happens to
    run mid-capture and finds an old model, engine or graph (seen once as ``cudaErrorStreamCaptureInvalidated`` at the first
    side-stream launch of a capture, in a long test session; never reproducible in isolation).  So: collect first, keep
    the collector off for the duration, and capture in "relaxed" mode -- the step body makes no legacy-stream or
    synchronising calls of its own, so nothing the relaxed mode admits can leak into the graph."""
    gc.collect()
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        with torch.cuda.graph(graph, stream=stream, capture_error_mode="relaxed"):
            yield
    finally:
        if was_enabled:
            gc.enable()
