"""Helper launched under torchrun by test_fused_engine_gpu.py: the flagship job on WORLD_SIZE processes over peer memory.
With fewer GPUs than processes several ranks share a device (Gloo bootstrap): the VMM fd export / import, the flag protocol and
the push / broadcast kernels are exercised across PROCESS boundaries even on a one-GPU box."""
import hashlib
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from draco_b200 import JobConfig  # noqa: E402
from draco_b200.parallel.trainer import Trainer, init_distributed  # noqa: E402


def main():
    kw = json.loads(os.environ.get("MP_EQUIV_CFG", "{}"))
    rank, world, local = init_distributed(kw.get("transport", "nvl"))
    ngpu = torch.cuda.device_count()
    base = dict(network="ResNet18", dataset="Cifar10", approach="maj_vote", mode="maj_vote", group_size=3, worker_fail=1,
                err_mode="rev_grad", batch_size=8, num_workers=7, max_steps=12, transport="nvl", lr=0.02, momentum=0.9,
                synthetic_size=256, eval_freq=10 ** 6, compress_grad="None", dtype="bf16", cuda_graphs=True,
                spin_timeout_s=120.0 if world > ngpu else 20.0)
    base.update(kw)
    steps = int(os.environ.get("MP_EQUIV_STEPS", "6"))
    cfg = JobConfig(**base)
    t = Trainer(cfg, rank=rank, world=world, device=torch.device("cuda", local % ngpu), quiet=True)
    losses = []
    for _ in range(steps):
        m = t.train_step()
        losses.append(m.get("loss") if m else None)
    t.synchronize()
    dist.barrier()
    p = t.engine.master_params()
    rec = {"rank": rank, "param_sum": p.double().sum().item(), "sha": hashlib.sha256(p.cpu().numpy().tobytes()).hexdigest(),
           "loss_first": losses[0], "loss_last": losses[-1]}
    recs = [None] * world
    dist.all_gather_object(recs, rec)
    if rank == 0 and os.environ.get("MP_EQUIV_OUT"):
        torch.save({"params": p.detach().cpu(), "losses": losses}, os.environ["MP_EQUIV_OUT"])
    if rank == 0:
        print(json.dumps({"param_sum": recs[0]["param_sum"], "worker_param_sum": recs[-1]["param_sum"],
                          "sha": [r["sha"] for r in recs], "losses": [[r["loss_first"], r["loss_last"]] for r in recs],
                          "multicast": bool(getattr(t.engine, "mc_params", None)), "placement": t.engine.place.describe(),
                          "gpus": ngpu, "world": world}), flush=True)
    t.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
