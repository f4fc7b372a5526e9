"""Patience-based stop criterion (reference: utils/customized.py:16-50)."""
from __future__ import annotations

from typing import Optional


class EarlyStopping:
    """Stop when a monitored metric has not improved for ``patience`` calls.

    ``stopper(metric)`` treats the metric as a loss (lower is better);
    ``stopper(metric, loss=False)`` treats it as a score.  A call counts as "no
    improvement" when the new score is below ``best + delta`` -- note an exactly
    equal score with ``delta == 0`` *is* an improvement, as in the reference.
    ``early_stop`` latches True once the counter reaches ``patience`` and the
    word ``EarlyStopping`` is printed when ``verbose``.

    Extension: :meth:`state_dict` / :meth:`load_state_dict` for resumable runs,
    and :func:`agree` so that every rank takes rank 0's decision instead of
    relying on bit-identical validation losses.
    """

    def __init__(self, patience: int = 100, verbose: bool = True, delta: float = 0):
        self.patience = patience
        self.verbose = verbose
        self.delta = delta
        self.counter = 0
        self.best_score: Optional[float] = None
        self.early_stop = False

    def __call__(self, metrics: float, loss: bool = True) -> None:
        score = -metrics if loss else metrics
        if self.best_score is not None and score < self.best_score + self.delta:
            self.counter += 1
            if self.counter >= self.patience:
                self.early_stop = True
                if self.verbose:
                    print("EarlyStopping")
            return
        self.best_score = score
        self.counter = 0

    # -- extensions -----------------------------------------------------
    def state_dict(self) -> dict:
        return {"counter": self.counter, "best_score": self.best_score,
                "early_stop": self.early_stop}

    def load_state_dict(self, state: dict) -> None:
        self.counter = int(state["counter"])
        self.best_score = state["best_score"]
        self.early_stop = bool(state["early_stop"])

    def agree(self, group=None) -> bool:
        """Broadcast rank 0's decision so no rank can leave the loop alone."""
        import torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return self.early_stop
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        flag = torch.tensor([int(self.early_stop)], dtype=torch.int32, device=dev)
        dist.broadcast(flag, src=0, group=group)
        self.early_stop = bool(flag.item())
        return self.early_stop
