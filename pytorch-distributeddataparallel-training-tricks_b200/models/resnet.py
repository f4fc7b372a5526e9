"""ResNet-50 (BASELINE.json config 4: random-init, 3x224x224 synthetic, DDP bf16).

The reference ships only Toy_Net; this family exists because the baseline spec
names it as the configuration that stresses gradient all-reduce overlap
(~102 MB of fp32 gradients in 25 MB buckets).  Standard v1.5 bottleneck layout
(stride on the 3x3), zero-initialised last BN gamma per block.  The pointwise
(1x1) convolutions -- 36 of the 53, ~51 % of the multiply-accumulates -- are
:class:`ddpt_b200.ops.conv_modules.TensorCoreConv1x1`: on CUDA in bf16 they run on
the framework's own tcgen05 GEMM kernels (forward, data gradient, weight gradient);
the 7x7 stem and the 3x3 convolutions use cuDNN.  Bucketing, fused all-reduce,
Lookahead-SGD, loss scaling and the CUDA-graphed step (engine.GraphedTrainStep)
are the framework's own.
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn as nn

from ..ops.conv_modules import TensorCoreConv1x1


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, c_in: int, width: int, stride: int = 1):
        super().__init__()
        c_out = width * self.expansion
        self.conv1 = TensorCoreConv1x1(c_in, width, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.conv2 = nn.Conv2d(width, width, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.conv3 = TensorCoreConv1x1(width, c_out, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(c_out)
        nn.init.zeros_(self.bn3.weight)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = None
        if stride != 1 or c_in != c_out:
            self.downsample = nn.Sequential(TensorCoreConv1x1(c_in, c_out, 1, stride=stride, bias=False),
                                            nn.BatchNorm2d(c_out))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        identity = x if self.downsample is None else self.downsample(x)
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        return self.relu(out + identity)


class ResNet50(nn.Module):
    def __init__(self, num_classes: int = 1000, in_channels: int = 3):
        super().__init__()
        self.stem = nn.Sequential(nn.Conv2d(in_channels, 64, 7, stride=2, padding=3, bias=False),
                                  nn.BatchNorm2d(64), nn.ReLU(inplace=True),
                                  nn.MaxPool2d(3, stride=2, padding=1))
        blocks: List[nn.Module] = []
        c_in = 64
        for width, depth, stride in ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)):
            for i in range(depth):
                blocks.append(Bottleneck(c_in, width, stride if i == 0 else 1))
                c_in = width * Bottleneck.expansion
        self.layers = nn.Sequential(*blocks)
        self.head = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(c_in, num_classes))
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.head(self.layers(self.stem(x)))


def resnet50(num_classes: int = 1000) -> ResNet50:
    model = ResNet50(num_classes)
    from ..ops.conv_modules import use_tensor_core_pointwise_convs
    use_tensor_core_pointwise_convs(model)       # records each layer's parameter index (bf16 shadow lookup)
    return model
