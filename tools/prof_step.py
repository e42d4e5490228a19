"""Small driver for ncu: a few eager steps of the flagship job packed on one GPU (so every step-path kernel launches)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from draco_b200 import JobConfig  # noqa: E402
from draco_b200.parallel.trainer import Trainer  # noqa: E402

approach = sys.argv[1] if len(sys.argv) > 1 else "maj_vote"
kw = dict(approach="maj_vote", mode="maj_vote", group_size=3, worker_fail=3, err_mode="rev_grad")
if approach == "cyclic":
    kw = dict(approach="cyclic", worker_fail=1, err_mode="constant")
elif approach in ("krum", "geometric_median"):
    kw = dict(approach="baseline", mode=approach, worker_fail=2, err_mode="rev_grad")
cfg = JobConfig(network="ResNet18", dataset="Cifar10", batch_size=128, num_workers=7, max_steps=16, transport="nvl", dtype="bf16",
                cuda_graphs=False, overlap_push=os.environ.get("OVERLAP", "0") == "1", compress_grad="None", synthetic_size=2048, eval_freq=10 ** 9, lr=0.01, momentum=0.9, **kw)
t = Trainer(cfg, rank=0, world=1, device=torch.device("cuda", 0), quiet=True)
for _ in range(int(os.environ.get("STEPS", "4"))):
    t.train_step()
t.close()
print("done")
