"""Host cost of one CUDA-graph replay of the flagship step vs its device time (is the step launch-bound?)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from draco_b200 import JobConfig  # noqa: E402
from draco_b200.parallel.trainer import Trainer  # noqa: E402

ws = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = JobConfig(network="ResNet18", dataset="Cifar10", approach="maj_vote", mode="maj_vote", batch_size=128, num_workers=7,
                group_size=3, worker_fail=3, err_mode="rev_grad", lr=0.01, momentum=0.9, max_steps=400, eval_freq=10 ** 9,
                transport="nvl", dtype="bf16", cuda_graphs=True, compress_grad="None", synthetic_size=8192, log_interval=10 ** 9,
                worker_streams=ws, data_on_device=True)
t = Trainer(cfg, rank=0, world=1, device=torch.device("cuda", 0), quiet=True)
eng = t.engine
for _ in range(6):
    t.train_step_async()
torch.cuda.synchronize()
g = eng.graph
assert g is not None
host = []
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20):
    t0 = time.perf_counter()
    g.replay()
    host.append(time.perf_counter() - t0)
e.record()
torch.cuda.synchronize()
dev = s.elapsed_time(e) / 20
# one replay alone, device idle before and after: pure device time of a step
lone = []
for _ in range(5):
    torch.cuda.synchronize()
    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
    lone.append(s.elapsed_time(e))
print(f"worker_streams={ws} kernels/step={eng.kernels_per_step} host replay ms: first {host[0]*1e3:.3f} median {sorted(host)[10]*1e3:.3f} "
      f"max {max(host)*1e3:.3f}; back-to-back device ms/step {dev:.3f}; lone replay ms {sorted(lone)[2]:.3f}")
