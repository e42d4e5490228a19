"""CIFAR VGG family.

Architecture parity with the reference (src/model_ops/vgg.py:15-108): torchvision-style feature stacks
(configs A/B/D/E = VGG-11/13/16/19, optional BatchNorm), followed by the small CIFAR classifier
Dropout -> Linear(512,512) -> ReLU -> Dropout -> Linear(512,512) -> ReLU -> Linear(512,10); conv weights
initialised N(0, sqrt(2/(k*k*Cout))) with zero bias (vgg.py:33-38).  vgg11_bn: 38 tensors / 9,756,426
parameters; vgg16_bn: 58 / 15,253,578.

Dropout makes replicas diverge unless every holder of a batch uses the same mask: the classifier uses
``ops.dropout.ReplicaDropout``, whose mask is a counter-based hash of (seed, step read from device memory, batch id,
layer, element) -- identical for every holder of a batch, also inside a replayed CUDA graph.  That is what makes VGG
usable under the repetition and cyclic codes (the reference only runs VGG in the baseline approach).
"""
from __future__ import annotations

import math
from typing import List, Union

import torch
from torch import nn

from ..ops.conv import Conv2d
from ..ops.dropout import ReplicaDropout
from ..ops.linear import Linear
from ..ops.norm import FusedBatchNorm2d
from .split import make_split

CFG = {
    "A": [64, "M", 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"],
    "B": [64, 64, "M", 128, 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"],
    "D": [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"],
    "E": [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"],
}


class _Features(nn.Sequential):
    """``nn.Sequential`` (same numbering / state_dict keys) that hands each convolution the BatchNorm that follows it, so that
    the convolution epilogue produces the batch statistics (ops/conv.py: ``Conv2d.forward(x, bn=...)``)."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        mods = list(self)
        for i, m in enumerate(mods):
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            if isinstance(m, Conv2d) and isinstance(nxt, FusedBatchNorm2d):
                x = m(x, bn=nxt)
            else:
                x = m(x)
        return x


def make_layers(cfg: List[Union[int, str]], batch_norm: bool = False) -> nn.Sequential:
    layers: List[nn.Module] = []
    cin = 3
    for v in cfg:
        if v == "M":
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
            continue
        layers.append(Conv2d(cin, int(v), kernel_size=3, padding=1))
        if batch_norm:
            # BN + ReLU fused (ops/norm.py); the Identity keeps torchvision's module numbering / state_dict keys
            layers.append(FusedBatchNorm2d(int(v), relu=True))
            layers.append(nn.Identity())
        else:
            layers.append(nn.ReLU(inplace=True))
        cin = int(v)
    return _Features(*layers)


class VGG(nn.Module):
    num_classes = 10
    input_shape = (3, 32, 32)

    def __init__(self, features: nn.Sequential, num_classes: int = 10) -> None:
        super().__init__()
        self.features = features
        self.classifier = nn.Sequential(
            ReplicaDropout(salt=1), Linear(512, 512), nn.ReLU(True),
            ReplicaDropout(salt=2), Linear(512, 512), nn.ReLU(True),
            Linear(512, num_classes),
        )
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                fan = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                nn.init.normal_(m.weight, 0.0, math.sqrt(2.0 / fan))
                nn.init.zeros_(m.bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.classifier(self.features(x).flatten(1))

    def name(self) -> str:
        return "vgg"


VGGSplit = make_split(VGG, "VGGSplit")


def _factory(cfg: str, bn: bool, name: str):
    def make(num_classes: int = 10):
        # every zoo model carries the layer-wise backward drivers the worker runtime uses (models/split.py)
        return VGGSplit(make_layers(CFG[cfg], batch_norm=bn), num_classes=num_classes)
    make.__name__ = name
    return make


vgg11, vgg11_bn = _factory("A", False, "vgg11"), _factory("A", True, "vgg11_bn")
vgg13, vgg13_bn = _factory("B", False, "vgg13"), _factory("B", True, "vgg13_bn")
vgg16, vgg16_bn = _factory("D", False, "vgg16"), _factory("D", True, "vgg16_bn")
vgg19, vgg19_bn = _factory("E", False, "vgg19"), _factory("E", True, "vgg19_bn")
