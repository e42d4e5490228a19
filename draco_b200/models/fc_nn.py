"""Fully connected MNIST net.

Architecture parity with the reference (src/model_ops/fc_nn.py:21-39): 784 -> 800 -> ReLU -> 500 -> ReLU ->
10 -> **Sigmoid**, and the sigmoid output is what is fed to CrossEntropyLoss (a reference quirk that
is part of the model).  6 parameter tensors / 1,033,510 parameters.
"""
from __future__ import annotations

import torch
from torch import nn

from ..ops.linear import Linear
from .split import make_split


class FC_NN(nn.Module):
    num_classes = 10
    input_shape = (1, 28, 28)

    def __init__(self) -> None:
        super().__init__()
        self.fc1 = Linear(784, 800)
        self.fc2 = Linear(800, 500)
        self.fc3 = Linear(500, 10)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = x.reshape(x.shape[0], -1)
        x = torch.relu(self.fc1(x))
        x = torch.relu(self.fc2(x))
        return torch.sigmoid(self.fc3(x))

    def name(self) -> str:
        return "fc_nn"


FC_NN_Split = make_split(FC_NN, "FC_NN_Split")
