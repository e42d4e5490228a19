"""``distributed_evaluator`` -- a separate process that follows a training run through its checkpoints.

Reference: src/distributed_evaluator.py:40-158 (+ src/evaluate_pytorch.sh): polls ``model_dir/model_step_{k*eval_freq}``
every 10 s, loads it, prints test loss / Prec@1 / Prec@5.  (The reference script has undefined names and expects a
state_dict while the master pickles whole modules; this one reads the checkpoint format of utils/checkpoint.py.)

    python -m draco_b200.cli.distributed_evaluator --eval-batch-size 1000 --eval-freq 50 --network ResNet18 \
           --dataset Cifar10 --model-dir output/models/
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import torch
import torch.nn.functional as F

from ..data import load_dataset
from ..models import build_model
from ..parallel.worker import accuracy
from ..utils.checkpoint import checkpoint_path, load_checkpoint, load_into_model


class DistributedEvaluator:
    def __init__(self, network: str, dataset: str, model_dir: str, eval_freq: int, eval_batch_size: int,
                 data_root: str = "./data", device: str = "auto", poll_s: float = 10.0, synthetic_size: int = 8192):
        self.model_dir, self.eval_freq, self.bs, self.poll_s = model_dir, eval_freq, eval_batch_size, poll_s
        self.device = torch.device("cuda" if (device == "auto" and torch.cuda.is_available()) else ("cpu" if device == "auto" else device))
        self.network = build_model(network).to(self.device)
        self.test_set = load_dataset(dataset, data_root, train=False, synthetic_size=synthetic_size)
        self.next_step = eval_freq

    @torch.no_grad()
    def _evaluate_model(self) -> dict:
        self.network.eval()
        ds = self.test_set
        mean, std = ds.mean.to(self.device), ds.std.to(self.device)
        tot, n = torch.zeros(3, device=self.device), 0
        for s in range(0, len(ds), self.bs):
            x = ((ds.images[s:s + self.bs].to(self.device).float() / 255.0) - mean) / std
            y = ds.labels[s:s + self.bs].to(self.device)
            out = self.network(x)
            p1, p5 = accuracy(out, y)
            tot += torch.stack([F.cross_entropy(out, y), p1, p5]) * y.shape[0]
            n += y.shape[0]
        loss, p1, p5 = (tot / max(n, 1)).tolist()
        return {"loss": loss, "prec1": p1, "prec5": p5}

    def evaluate_once(self, step: int) -> dict:
        blob = load_checkpoint(checkpoint_path(self.model_dir, step))
        load_into_model(blob, self.network)
        res = self._evaluate_model()
        print("Evaluator evaluating results on step {}: Test set: Average loss: {:.4f}, Prec@1: {:.3f} Prec@5: {:.3f}".format(
            step, res["loss"], res["prec1"], res["prec5"]), flush=True)
        return res

    def evaluate(self, max_evals: int = 0, timeout_s: float = 0.0) -> int:
        done, t0 = 0, time.time()
        while True:
            path = checkpoint_path(self.model_dir, self.next_step)
            if os.path.isfile(path):
                self.evaluate_once(self.next_step)
                self.next_step += self.eval_freq
                done += 1
                t0 = time.time()
                if max_evals and done >= max_evals:
                    return done
            else:
                if timeout_s and time.time() - t0 > timeout_s:
                    return done
                time.sleep(self.poll_s)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="draco_b200 evaluator")
    ap.add_argument("--eval-batch-size", type=int, default=1000)
    ap.add_argument("--eval-freq", type=int, default=50)
    ap.add_argument("--model-dir", type=str, default="output/models/")
    ap.add_argument("--dataset", type=str, default="MNIST")
    ap.add_argument("--network", type=str, default="LeNet")
    ap.add_argument("--data-root", type=str, default="./data")
    ap.add_argument("--device", type=str, default="auto")
    ap.add_argument("--poll-s", type=float, default=10.0)
    ap.add_argument("--max-evals", type=int, default=0)
    ap.add_argument("--timeout-s", type=float, default=0.0)
    a = ap.parse_args(argv)
    ev = DistributedEvaluator(a.network, a.dataset, a.model_dir, a.eval_freq, a.eval_batch_size, a.data_root, a.device, a.poll_s)
    ev.evaluate(a.max_evals, a.timeout_s)
    return 0


if __name__ == "__main__":
    sys.exit(main())
