"""Model zoo (the reference's ``model_ops``): LeNet, FC, ResNet-18..152, VGG-11..19 (+ ``*Split`` drivers).

``build_model(name)`` maps the reference's ``--network`` values (src/distributed_nn.py:45,
src/master/baseline_master.py:30-52) to constructors.  Unlike the reference every approach
(baseline / maj_vote / cyclic) supports every network.
"""
from __future__ import annotations

from typing import Callable, Dict

from torch import nn

from .fc_nn import FC_NN, FC_NN_Split
from .lenet import LeNet, LeNetSplit
from .resnet import (ResNet, ResNet18, ResNet34, ResNet50, ResNet101, ResNet152, ResNetSplit,
                     ResNetSplit18, ResNetSplit34, ResNetSplit50, ResNetSplit101, ResNetSplit152)
from .split import SplitBackwardMixin
from .vgg import (VGG, vgg11, vgg11_bn, vgg13, vgg13_bn, vgg16, vgg16_bn, vgg19, vgg19_bn)

_REGISTRY: Dict[str, Callable[..., nn.Module]] = {
    "LeNet": LeNetSplit,
    "FC": FC_NN_Split,
    "ResNet18": ResNetSplit18,
    "ResNet34": ResNetSplit34,
    "ResNet50": ResNetSplit50,
    "ResNet101": ResNetSplit101,
    "ResNet152": ResNetSplit152,
    # the reference's --network VGG11/13/16 instantiate the *_bn variants (baseline_master.py:47-52)
    "VGG11": vgg11_bn,
    "VGG13": vgg13_bn,
    "VGG16": vgg16_bn,
    "VGG19": vgg19_bn,
    "VGG11_plain": vgg11,
    "VGG13_plain": vgg13,
    "VGG16_plain": vgg16,
    "VGG19_plain": vgg19,
}

DATASET_OF = {"LeNet": "MNIST", "FC": "MNIST"}


def available_networks():
    return sorted(_REGISTRY)


def build_model(name: str, **kwargs) -> nn.Module:
    try:
        ctor = _REGISTRY[name]
    except KeyError:
        raise ValueError(f"unknown --network {name!r}; choose from {available_networks()}") from None
    return ctor(**kwargs)


__all__ = [n for n in dir() if not n.startswith("_")]
