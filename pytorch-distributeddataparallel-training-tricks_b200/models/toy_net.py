"""Toy_Net -- the reference's MNIST CNN (reference: utils/model.py:4-34).

Architecture (per image): 4 x [Conv3x3(valid) -> BatchNorm2d -> ReLU] with a
MaxPool2d(2) after the 2nd and 4th block, then Flatten -> Linear(8192,512) ->
BatchNorm1d -> ReLU -> Linear(512,10).  5 752 714 parameters.

Parity contract
---------------
* Sub-module names/indices (``conv.0 .. conv.13``, ``dense.0 .. dense.4``) and
  therefore the 37 ``state_dict`` keys are identical to the reference, so
  checkpoints are interchangeable in both directions.
* Modules are created in definition order with torch's stock initialisers, so
  ``same_seeds(s); Toy_Net()`` consumes the RNG stream exactly like the
  reference and produces bit-identical initial weights.

B200 path
---------
On a CUDA device, when the sm_100a extension is loaded, ``forward`` runs the
fused NHWC/bf16 pipeline in :mod:`ddpt_b200.ops.toynet_fused` (tcgen05 implicit
GEMM convolutions with BatchNorm statistics in the epilogue, a fused
normalise+ReLU(+pool) pass, tcgen05 linear layers).  On CPU -- and as the
numerical oracle in tests -- it is the plain ``nn.Sequential`` below.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.nn as nn

# (out_channels, pool_after) for the four conv blocks; spatial sizes 28->26->24->12->10->8->4
TOY_NET_CONV_PLAN: Tuple[Tuple[int, bool], ...] = ((64, False), (128, True), (256, False), (512, True))
_HIDDEN, _CLASSES = 512, 10


def _conv_stack(in_channels: int = 1) -> nn.Sequential:
    layers: List[nn.Module] = []
    c_in = in_channels
    for c_out, pool in TOY_NET_CONV_PLAN:
        layers += [nn.Conv2d(c_in, c_out, kernel_size=3), nn.BatchNorm2d(c_out), nn.ReLU(inplace=True)]
        if pool:
            layers.append(nn.MaxPool2d(2))
        c_in = c_out
    return nn.Sequential(*layers)


def _feature_dim(size: int = 28) -> int:
    c = 1
    for c_out, pool in TOY_NET_CONV_PLAN:
        size -= 2
        if pool:
            size //= 2
        c = c_out
    return c * size * size  # 512 * 4 * 4 = 8192


class Toy_Net(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.conv = _conv_stack(1)
        self.dense = nn.Sequential(
            nn.Flatten(),
            nn.Linear(_feature_dim(), _HIDDEN),
            nn.BatchNorm1d(_HIDDEN),
            nn.ReLU(inplace=True),
            nn.Linear(_HIDDEN, _CLASSES),
        )
        self._fused = None  # set by ops.toynet_fused.attach(); never serialised

    def forward_reference(self, x: torch.Tensor) -> torch.Tensor:
        """Plain PyTorch graph -- the CPU path and the oracle for kernel tests."""
        if x.dtype == torch.uint8:
            x = x.float().div_(255.0)
        return self.dense(self.conv(x))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        fused = self._fused
        if fused is not None and x.is_cuda:
            return fused(x)
        return self.forward_reference(x)


def toy_net_flops_per_image() -> float:
    """Forward multiply-add FLOPs per image (2*MACs); 304.1 MFLOP."""
    flops, size, c_in = 0.0, 28, 1
    for c_out, pool in TOY_NET_CONV_PLAN:
        size -= 2
        flops += 2.0 * size * size * c_out * c_in * 9
        if pool:
            size //= 2
        c_in = c_out
    flops += 2.0 * _feature_dim() * _HIDDEN + 2.0 * _HIDDEN * _CLASSES
    return flops
