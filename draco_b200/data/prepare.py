"""Dataset preparation (reference: src/datasets/data_prepare.py:1-28).

With network access this downloads MNIST / CIFAR-10 through torchvision into ``<root>/mnist_data`` and
``<root>/cifar10_data`` (the directories ``load_dataset`` looks for).  Offline it materialises the synthetic stand-ins as
``.pt`` files so every rank maps the same bytes.
"""
from __future__ import annotations

import argparse
import os

import torch

from . import synthetic_dataset


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--root", default="./data")
    ap.add_argument("--download", action="store_true", help="try torchvision downloads (needs network)")
    ap.add_argument("--synthetic-size", type=int, default=8192)
    a = ap.parse_args(argv)
    os.makedirs(a.root, exist_ok=True)
    if a.download:
        from torchvision import datasets
        for train in (True, False):
            datasets.MNIST(os.path.join(a.root, "mnist_data"), train=train, download=True)
            datasets.CIFAR10(os.path.join(a.root, "cifar10_data"), train=train, download=True)
        print("downloaded MNIST and CIFAR-10")
        return 0
    for name in ("MNIST", "Cifar10"):
        ds = synthetic_dataset(name, a.synthetic_size)
        path = os.path.join(a.root, f"synthetic_{name}.pt")
        torch.save({"images": ds.images, "labels": ds.labels}, path)
        print("wrote", path, tuple(ds.images.shape))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
