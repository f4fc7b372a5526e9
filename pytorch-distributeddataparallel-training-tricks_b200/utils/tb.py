"""TensorBoard scalars with the reference's on-disk layout (utils/train.py:64-65,95-98).

Root run holds tag ``LR``; ``add_scalars("Loss", {"train","valid"})`` creates the
sub-runs ``Loss_train`` / ``Loss_valid`` (tag ``Loss``) and likewise for ``Acc``.
Falls back to a JSON-lines file when tensorboard is not importable.
"""
from __future__ import annotations

import json
import os
import time


class ScalarLogger:
    def __init__(self, log_dir: str):
        self.log_dir = log_dir
        os.makedirs(log_dir, exist_ok=True)
        self._writer = None
        self._jsonl = None
        try:
            from torch.utils.tensorboard import SummaryWriter
            self._writer = SummaryWriter(log_dir)
        except Exception:  # tensorboard missing: keep the numbers anyway
            self._jsonl = open(os.path.join(log_dir, "scalars.jsonl"), "a")

    def epoch(self, epoch: int, lr: float, train_loss: float, valid_loss: float,
              train_acc: float, valid_acc: float) -> None:
        if self._writer is not None:
            self._writer.add_scalar("LR", lr, epoch)
            self._writer.add_scalars("Loss", {"train": train_loss, "valid": valid_loss}, epoch)
            self._writer.add_scalars("Acc", {"train": train_acc, "valid": valid_acc}, epoch)
        else:
            self._jsonl.write(json.dumps({"t": time.time(), "epoch": epoch, "LR": lr,
                                          "Loss": {"train": train_loss, "valid": valid_loss},
                                          "Acc": {"train": train_acc, "valid": valid_acc}}) + "\n")
            self._jsonl.flush()

    def close(self) -> None:
        if self._writer is not None:
            self._writer.close()
        if self._jsonl is not None:
            self._jsonl.close()
