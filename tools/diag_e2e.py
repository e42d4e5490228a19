"""Host timeline of the pipelined e2e loop (where do the milliseconds between graph replays go?)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from draco_b200 import JobConfig  # noqa: E402
from draco_b200.parallel.trainer import Trainer  # noqa: E402

cfg = JobConfig(network="ResNet18", dataset="Cifar10", approach="maj_vote", mode="maj_vote", batch_size=128, num_workers=7,
                group_size=3, worker_fail=3, err_mode="rev_grad", lr=0.01, momentum=0.9, max_steps=400, eval_freq=10 ** 9,
                transport="nvl", dtype="bf16", cuda_graphs=True, compress_grad="None", synthetic_size=8192, log_interval=10 ** 9,
                data_on_device=False)
t = Trainer(cfg, rank=0, world=1, device=torch.device("cuda", 0), quiet=True)
eng = t.engine
for _ in range(6):
    t.train_step_pipelined()
t.drain()
torch.cuda.synchronize()
rows = []
pending = None
T = time.perf_counter
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t_all = T()
s.record()
for i in range(20):
    t0 = T()
    eng.graph.replay()
    t1 = T()
    eng.step += 1
    eng._stage(eng.step)
    eng._staged_step = eng.step
    t2 = T()
    h = eng.enqueue_metrics_read()
    t3 = T()
    if pending is not None:
        eng.resolve_metrics(pending)
    pending = h
    t4 = T()
    rows.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
eng.resolve_metrics(pending)
e.record()
torch.cuda.synchronize()
wall = (T() - t_all) / 20 * 1e3
import statistics as st
try:
    print("cgroup cpu.max:", open("/sys/fs/cgroup/cpu.max").read().strip(), "| cpus:", len(os.sched_getaffinity(0)),
          "| torch threads:", torch.get_num_threads())
except OSError:
    pass
names = ("replay", "stage(next)", "enqueue_metrics", "resolve(prev)")
print("wall ms/step %.3f device ms/step %.3f" % (wall, s.elapsed_time(e) / 20))
print("per-step total ms:", " ".join("%.1f" % (sum(r) * 1e3) for r in rows), "| engine step at loop start:", eng.step - 20)
for j, n in enumerate(names):
    v = [r[j] * 1e3 for r in rows[2:]]
    print("  %-16s median %.3f max %.3f" % (n, st.median(v), max(v)))
