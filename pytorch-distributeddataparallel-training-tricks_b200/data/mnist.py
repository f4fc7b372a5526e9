"""MNIST as two dense arrays (reference: utils/train.py:25,29 via torchvision).

The reference materialises every sample through PIL + ``ToTensor`` in the
training process (1024 Python calls per step).  Here the whole
split is one ``uint8 [N,28,28]`` tensor plus one ``int64 [N]`` tensor -- 47 MB for
the training split -- parsed straight from the IDX files; ``x/255`` happens on the
device inside the first kernel of the network (or vectorised on the host for the
CPU path).  Numerically identical to ``ToTensor`` (uint8 -> float32 / 255).

Search order under ``root`` (the reference's ``-d`` flag): ``MNIST/raw/*-ubyte``,
``MNIST/raw/*-ubyte.gz``, ``MNIST/processed/{training,test}.pt``.
"""
from __future__ import annotations

import gzip
import os
import struct
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

_FILES = {True: ("train-images-idx3-ubyte", "train-labels-idx1-ubyte"),
          False: ("t10k-images-idx3-ubyte", "t10k-labels-idx1-ubyte")}
_IDX_DTYPES = {0x08: np.uint8, 0x09: np.int8, 0x0B: ">i2", 0x0C: ">i4", 0x0D: ">f4", 0x0E: ">f8"}


@dataclass
class MNISTArrays:
    images: torch.Tensor  # uint8 [N, 28, 28]
    labels: torch.Tensor  # int64 [N]
    source: str

    def __len__(self) -> int:
        return int(self.images.shape[0])


def _read_idx(path: str) -> np.ndarray:
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as f:
        raw = f.read()
    zero, dtype_code, ndim = struct.unpack_from(">HBB", raw, 0)
    if zero != 0 or dtype_code not in _IDX_DTYPES:
        raise ValueError(f"{path}: not an IDX file")
    dims = struct.unpack_from(">" + "I" * ndim, raw, 4)
    data = np.frombuffer(raw, dtype=_IDX_DTYPES[dtype_code], offset=4 + 4 * ndim)
    return data.reshape(dims)


def _first_existing(*paths: str) -> Optional[str]:
    for p in paths:
        if os.path.exists(p):
            return p
    return None


def load_mnist(root: str, train: bool = True) -> MNISTArrays:
    """Load one MNIST split from ``root`` without torchvision/PIL."""
    raw_dir = os.path.join(root, "MNIST", "raw")
    img_name, lbl_name = _FILES[train]
    img_path = _first_existing(os.path.join(raw_dir, img_name), os.path.join(raw_dir, img_name + ".gz"))
    lbl_path = _first_existing(os.path.join(raw_dir, lbl_name), os.path.join(raw_dir, lbl_name + ".gz"))
    if img_path and lbl_path:
        images = torch.from_numpy(np.array(_read_idx(img_path), copy=True))
        labels = torch.from_numpy(_read_idx(lbl_path).astype(np.int64))
        return MNISTArrays(images, labels, img_path)
    processed = os.path.join(root, "MNIST", "processed", "training.pt" if train else "test.pt")
    if os.path.exists(processed):
        images, labels = torch.load(processed, map_location="cpu", weights_only=True)
        return MNISTArrays(images.to(torch.uint8).contiguous(), labels.to(torch.int64), processed)
    raise FileNotFoundError(
        f"MNIST not found under {root!r} (looked for MNIST/raw/{img_name}[.gz] and "
        f"MNIST/processed/*.pt); pass --synthetic to train on generated data")


def synthetic_mnist(n: int, seed: int = 0, num_classes: int = 10) -> MNISTArrays:
    """Deterministic MNIST-shaped data: class-dependent blobs + noise (learnable)."""
    g = torch.Generator().manual_seed(seed)
    labels = torch.randint(0, num_classes, (n,), generator=g, dtype=torch.int64)
    protos = torch.rand((num_classes, 28, 28), generator=g)
    noise = torch.rand((n, 28, 28), generator=g)
    images = ((0.6 * protos[labels] + 0.4 * noise) * 255.0).to(torch.uint8)
    return MNISTArrays(images, labels, f"synthetic(seed={seed})")
