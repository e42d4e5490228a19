"""Helper launched under torchrun by test_fused_engine_gpu.py: the flagship job on WORLD_SIZE GPUs over peer memory."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from draco_b200 import JobConfig  # noqa: E402
from draco_b200.parallel.trainer import Trainer, init_distributed  # noqa: E402


def main():
    rank, world, local = init_distributed("nvl")
    cfg = JobConfig(network="ResNet18", dataset="Cifar10", approach="maj_vote", mode="maj_vote", group_size=3, worker_fail=1,
                    err_mode="rev_grad", batch_size=8, num_workers=7, max_steps=12, transport="nvl", lr=0.02, momentum=0.9,
                    synthetic_size=256, eval_freq=10 ** 6, compress_grad="None", dtype="bf16", cuda_graphs=True)
    t = Trainer(cfg, rank=rank, world=world, device=torch.device("cuda", local), quiet=True)
    for _ in range(6):
        t.train_step()
    t.synchronize()
    dist.barrier()
    s = t.engine.master_params().double().sum().item()
    sums = [None] * world
    dist.all_gather_object(sums, s)
    if rank == 0:
        print(json.dumps({"param_sum": sums[0], "worker_param_sum": sums[-1], "multicast": bool(t.engine.mc_params),
                          "placement": t.engine.place.describe()}), flush=True)
    t.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
