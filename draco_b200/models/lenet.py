"""LeNet for 1x28x28 inputs.

Architecture parity with the reference (src/model_ops/lenet.py:20-41): conv(1->20,k5) -> maxpool2 -> relu
-> conv(20->50,k5) -> maxpool2 -> relu -> fc(800->500) -> fc(500->10).  Note the reference's quirks that
are part of the architecture: pooling happens *before* the ReLU and there is no non-linearity between
fc1 and fc2.  8 parameter tensors / 431,080 parameters.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from ..ops.conv import Conv2d
from ..ops.linear import Linear
from .split import make_split


class LeNet(nn.Module):
    num_classes = 10
    input_shape = (1, 28, 28)

    def __init__(self) -> None:
        super().__init__()
        self.conv1 = Conv2d(1, 20, 5, 1)          # ops.conv.Conv2d: native kernels where the geometry is served
        self.conv2 = Conv2d(20, 50, 5, 1)
        self.fc1 = Linear(4 * 4 * 50, 500)
        self.fc2 = Linear(500, 10)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = F.relu(F.max_pool2d(self.conv1(x), 2, 2))
        x = F.relu(F.max_pool2d(self.conv2(x), 2, 2))
        x = x.reshape(x.shape[0], -1)
        return self.fc2(self.fc1(x))

    def name(self) -> str:
        return "lenet"


LeNetSplit = make_split(LeNet, "LeNetSplit")
