"""One pass over a loader -- the hot loop (reference: utils/process.py:6-42).

Same signature and return values as the reference::

    loss, acc, lr = iterate_loader(loader, model, loss_function, local_rank, optimizer, training=True)
    loss, acc     = iterate_loader(loader, model, loss_function, local_rank, None, training=False)

and the same arithmetic per batch, including quirk Q1 (``CrossEntropyLoss`` mean,
then divided by the batch size *again*, so gradients carry an extra 1/B and the
logged loss is CE/B).  What changed is where the work happens:

* loss and correct-count are accumulated **on the device** and read once at the
  end of the pass (the reference did two ``.item()`` host syncs per step);
* H2D copies are ``non_blocking`` from pinned memory;
* on CUDA, when the model/optimizer pair has a fused step engine attached
  (``ops.step_engine``), the whole step -- forward, loss, scaled backward with
  bucketed fused all-reduce, unscale/overflow, Lookahead-SGD -- is one CUDA-graph
  replay per batch with zero host synchronisation.

``local_rank`` is only the device index here; "am I rank 0" decisions use the
process group's rank (the reference conflated the two, quirk Q8).
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import amp


def _device_of(model: torch.nn.Module, local_rank) -> torch.device:
    try:
        return next(model.parameters()).device
    except StopIteration:
        return torch.device("cuda", local_rank) if torch.cuda.is_available() else torch.device("cpu")


def _progress(loader, enabled: bool):
    if not enabled:
        return loader
    try:
        from tqdm import tqdm
        return tqdm(loader, total=len(loader))
    except Exception:
        return loader


def iterate_loader(loader, model: torch.nn.Module, loss_function, local_rank: int,
                   apex_optimizer: Optional[torch.optim.Optimizer], training: bool = False,
                   progress: Optional[bool] = None):
    device = _device_of(model, local_rank)
    if progress is None:
        progress = not getattr(loader, "quiet", False)
    engine = getattr(model, "_ddpt_step_engine", None) if device.type == "cuda" else None

    if engine is not None and engine.accepts(model, loss_function, apex_optimizer, training):
        # fused path: one CUDA-graph replay per batch, metrics stay on the device until the end of the pass
        engine.begin_pass()
        for image, target in _progress(loader, progress):
            engine.run(image, target, training)
        loss_total, correct_total, seen = engine.end_pass()
        mean_loss, accuracy = loss_total / max(seen, 1), correct_total / max(seen, 1)
        if training:
            return mean_loss, accuracy, apex_optimizer.param_groups[0]["lr"]
        return mean_loss, accuracy

    loss_sum = torch.zeros((), dtype=torch.float64 if device.type == "cpu" else torch.float32, device=device)
    correct = torch.zeros((), dtype=torch.int64, device=device)
    seen = 0

    for image, target in _progress(loader, progress):
        if image.device != device:
            image = image.to(device, non_blocking=True)
        if target.device != device or target.dtype != torch.long:
            target = target.to(device, dtype=torch.long, non_blocking=True)
        batch = image.shape[0]

        if training:
            apex_optimizer.zero_grad()
        outputs = model(image)
        predict = torch.argmax(outputs, dim=1)
        batch_loss = loss_function(outputs.float(), target)
        batch_loss = batch_loss / len(outputs)          # Q1: second normalisation, as the reference
        if training:
            with amp.scale_loss(batch_loss, apex_optimizer) as scaled_loss:
                scaled_loss.backward()
            apex_optimizer.step()
        loss_sum += batch_loss.detach() * batch
        correct += (predict == target).sum()
        seen += batch

    mean_loss = float(loss_sum.item()) / max(seen, 1)
    accuracy = float(correct.item()) / max(seen, 1)
    if training:
        return mean_loss, accuracy, apex_optimizer.param_groups[0]["lr"]
    return mean_loss, accuracy
