"""Checkpointing.

* ``save_weights`` reproduces the reference artefact (utils/train.py:112-115):
  ``{model_path}/{exp_name}.pt`` holding the *unwrapped* module's ``state_dict``
  -- 37 keys for Toy_Net, fp32 weights + BN buffers, no ``module.`` prefix.
  Tensors are made dense/contiguous first so the file loads into the reference
  ``Toy_Net`` with ``strict=True`` regardless of our internal memory format.
* ``save_resume_state`` is the superset the reference lacks: SGD
  momentum, Lookahead slow weights + counter, both schedulers, early-stopping
  state, loss scaler, epoch, best accuracy and RNG state, in a sidecar file.
"""
from __future__ import annotations

import os
import random
from collections import OrderedDict
from typing import Any, Dict, Optional

import numpy as np
import torch


def _unwrap(model: torch.nn.Module) -> torch.nn.Module:
    return model.module if hasattr(model, "module") and isinstance(model.module, torch.nn.Module) else model


def dense_state_dict(model: torch.nn.Module) -> "OrderedDict[str, torch.Tensor]":
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, value in _unwrap(model).state_dict().items():
        out[key] = value.detach().clone(memory_format=torch.contiguous_format)
    return out


def save_weights(model: torch.nn.Module, path: str) -> None:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    tmp = f"{path}.tmp.{os.getpid()}"
    torch.save(dense_state_dict(model), tmp)
    os.replace(tmp, path)  # atomic: a killed run never leaves a torn checkpoint


def load_weights(model: torch.nn.Module, path: str, strict: bool = True):
    state = torch.load(path, map_location="cpu", weights_only=True)
    return _unwrap(model).load_state_dict(state, strict=strict)


def _rng_state() -> Dict[str, Any]:
    state = {"python": random.getstate(), "numpy": np.random.get_state(),
             "torch": torch.get_rng_state()}
    if torch.cuda.is_available():
        state["cuda"] = torch.cuda.get_rng_state_all()
    return state


def _set_rng_state(state: Dict[str, Any]) -> None:
    random.setstate(state["python"])
    np.random.set_state(state["numpy"])
    torch.set_rng_state(state["torch"])
    if "cuda" in state and torch.cuda.is_available():
        try:
            torch.cuda.set_rng_state_all(state["cuda"])
        except Exception:
            pass  # device count changed between save and resume


def save_resume_state(path: str, *, model, optimizer, schedulers: Dict[str, Any],
                      early_stopping=None, scaler=None, epoch: int, best_valid_acc: float,
                      extra: Optional[dict] = None) -> None:
    payload = {
        "format": "ddpt_b200.resume.v1",
        "epoch": epoch,
        "best_valid_acc": best_valid_acc,
        "model": dense_state_dict(model),
        "optimizer": optimizer.state_dict(),
        "schedulers": {k: s.state_dict() for k, s in schedulers.items()},
        "early_stopping": early_stopping.state_dict() if early_stopping is not None else None,
        "scaler": scaler.state_dict() if scaler is not None else None,
        "rng": _rng_state(),
        "extra": extra or {},
    }
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    tmp = f"{path}.tmp.{os.getpid()}"
    torch.save(payload, tmp)
    os.replace(tmp, path)


def load_resume_state(path: str, *, model, optimizer, schedulers: Dict[str, Any],
                      early_stopping=None, scaler=None, restore_rng: bool = True) -> dict:
    payload = torch.load(path, map_location="cpu", weights_only=False)
    if payload.get("format") != "ddpt_b200.resume.v1":
        raise ValueError(f"{path}: not a ddpt_b200 resume file")
    _unwrap(model).load_state_dict(payload["model"], strict=True)
    optimizer.load_state_dict(payload["optimizer"])
    for key, sched in schedulers.items():
        sched.load_state_dict(payload["schedulers"][key])
    if early_stopping is not None and payload["early_stopping"] is not None:
        early_stopping.load_state_dict(payload["early_stopping"])
    if scaler is not None and payload["scaler"] is not None:
        scaler.load_state_dict(payload["scaler"])
    if restore_rng:
        _set_rng_state(payload["rng"])
    return {"epoch": payload["epoch"], "best_valid_acc": payload["best_valid_acc"],
            "extra": payload.get("extra", {})}
