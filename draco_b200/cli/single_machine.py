"""Single-machine trainer: the reference's ``NN_Trainer`` / ``single_machine.py`` (src/nn_ops/__init__.py:29-114,
src/single_machine.py:183-211) -- one process, one model, plain SGD, train + validate; exercises ``backward_single``.

    python -m draco_b200.cli.single_machine --network LeNet --dataset MNIST --max-steps 200
"""
from __future__ import annotations

import argparse
import sys

import torch
import torch.nn.functional as F

from ..data import load_dataset
from ..models import build_model
from ..parallel.worker import accuracy


class NN_Trainer:
    def __init__(self, **kwargs):
        self.batch_size = kwargs.get("batch_size", 128)
        self.lr = kwargs.get("learning_rate", 0.01)
        self.momentum = kwargs.get("momentum", 0.9)
        self.max_steps = kwargs.get("max_steps", 100)
        self.network_config = kwargs.get("network", "LeNet")
        self.dataset_name = kwargs.get("dataset", "MNIST")
        self.device = torch.device(kwargs.get("device", "cuda" if torch.cuda.is_available() else "cpu"))

    def build_model(self):
        torch.manual_seed(1)
        self.network = build_model(self.network_config).to(self.device)
        self.optimizer = torch.optim.SGD(self.network.parameters(), lr=self.lr, momentum=self.momentum)
        return self

    def _batch(self, ds, step):
        n = len(ds)
        idx = torch.arange(step * self.batch_size, (step + 1) * self.batch_size) % n
        x = ds.normalize(ds.images[idx]).to(self.device)
        return x, ds.labels[idx].to(self.device)

    def train_and_validate(self, train_set=None, test_set=None, log_every: int = 10):
        train_set = train_set or load_dataset(self.dataset_name, train=True)
        test_set = test_set or load_dataset(self.dataset_name, train=False)
        self.network.train()
        loss = None
        for step in range(self.max_steps):
            x, y = self._batch(train_set, step)
            self.optimizer.zero_grad()
            out = self.network(x)
            loss = F.cross_entropy(out, y)
            if hasattr(self.network, "backward_single"):
                self.network.backward_single(loss)
            else:
                loss.backward()
            self.optimizer.step()
            if step % log_every == 0:
                p1, p5 = accuracy(out.detach(), y)
                print("Step: {}, Loss: {:.4f}, Prec@1: {:.2f}, Prec@5: {:.2f}".format(step, loss.item(), p1.item(), p5.item()), flush=True)
        return self.validate(test_set), float(loss.item()) if loss is not None else float("nan")

    @torch.no_grad()
    def validate(self, test_set):
        self.network.eval()
        x = test_set.normalize(test_set.images[:1000]).to(self.device)
        y = test_set.labels[:1000].to(self.device)
        out = self.network(x)
        p1, p5 = accuracy(out, y)
        self.network.train()
        print("Test set: Prec@1: {:.2f} Prec@5: {:.2f}".format(p1.item(), p5.item()), flush=True)
        return p1.item()


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="draco_b200 single-machine trainer")
    ap.add_argument("--network", default="LeNet")
    ap.add_argument("--dataset", default="MNIST")
    ap.add_argument("--batch-size", type=int, default=128)
    ap.add_argument("--lr", type=float, default=0.01)
    ap.add_argument("--momentum", type=float, default=0.9)
    ap.add_argument("--max-steps", type=int, default=100)
    a = ap.parse_args(argv)
    NN_Trainer(batch_size=a.batch_size, learning_rate=a.lr, momentum=a.momentum, max_steps=a.max_steps,
               network=a.network, dataset=a.dataset).build_model().train_and_validate()
    return 0


if __name__ == "__main__":
    sys.exit(main())
