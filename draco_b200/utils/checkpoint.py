"""Checkpoint / resume.

Reference behaviour: the PS pickles the whole module to ``train_dir + "model_step_" + step`` every
``eval_freq`` steps (src/master/baseline_master.py:237-243), worker 1 saves a state_dict for ResNets
(src/worker/baseline_worker.py:298-302), and resume exists only for the baseline approach from a hard-coded path without
optimizer state (baseline_master.py:54-57).

Here one file per checkpoint, same naming (``<train_dir>model_step_<N>``), containing logical (NCHW) parameter tensors,
the momentum buffers, the step, the BatchNorm statistics of the evaluating worker (the PS never runs a forward pass, so
its own BN statistics would be meaningless -- the reason the reference refuses to save ResNets on the PS) and the job
config.  Resume works for every approach and restores optimizer state.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from ..config import JobConfig
from ..parallel.arena import ArenaLayout


def checkpoint_path(train_dir: str, step: int) -> str:
    return f"{train_dir}model_step_{step}"


def save_checkpoint(path: str, layout: ArenaLayout, params: torch.Tensor, momentum: Optional[torch.Tensor], step: int,
                    cfg: JobConfig, buffers: Optional[Dict[str, torch.Tensor]] = None,
                    opt_state: Optional[Dict[str, torch.Tensor]] = None) -> None:
    """``momentum``: SGD momentum buffer / Adam first moment arena; ``opt_state``: further optimizer arenas by name
    (Adam: ``exp_avg_sq``, AMSGrad: ``max_exp_avg_sq``) -- all saved per parameter tensor."""
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    state = {layout.specs[i].name: layout.view(params, i).detach().float().cpu().contiguous().clone()
             for i in range(layout.ntensors)}
    mom = None
    if momentum is not None:
        mom = {layout.specs[i].name: layout.view(momentum, i).detach().float().cpu().contiguous().clone()
               for i in range(layout.ntensors)}
    extra = {name: {layout.specs[i].name: layout.view(arena, i).detach().float().cpu().contiguous().clone()
                    for i in range(layout.ntensors)} for name, arena in (opt_state or {}).items() if arena is not None}
    blob = {"format": "draco_b200/1", "step": int(step), "state_dict": state, "momentum": mom, "opt_state": extra,
            "buffers": {k: v.detach().cpu().clone() for k, v in (buffers or {}).items()}, "config": cfg.to_dict()}
    tmp = path + ".tmp"
    torch.save(blob, tmp)
    os.replace(tmp, path)


def load_checkpoint(path: str) -> dict:
    blob = torch.load(path, map_location="cpu", weights_only=False)
    if blob.get("format") != "draco_b200/1":
        raise ValueError(f"{path}: not a draco_b200 checkpoint")
    return blob


def restore_into(blob: dict, layout: ArenaLayout, params: torch.Tensor, momentum: Optional[torch.Tensor],
                 opt_state: Optional[Dict[str, torch.Tensor]] = None) -> int:
    with torch.no_grad():
        for i, spec in enumerate(layout.specs):
            layout.view(params, i).copy_(blob["state_dict"][spec.name])
            if momentum is not None and blob.get("momentum"):
                layout.view(momentum, i).copy_(blob["momentum"][spec.name])
            for name, arena in (opt_state or {}).items():
                saved = (blob.get("opt_state") or {}).get(name)
                if arena is not None and saved:
                    layout.view(arena, i).copy_(saved[spec.name])
    return int(blob["step"])


def model_buffers(model: torch.nn.Module) -> Dict[str, torch.Tensor]:
    return {k: v for k, v in model.named_buffers()}


def load_into_model(blob: dict, model: torch.nn.Module) -> None:
    """Load a checkpoint into a plain (unbound) model, e.g. in the evaluator."""
    sd = dict(blob["state_dict"])
    sd.update(blob.get("buffers") or {})
    model.load_state_dict(sd, strict=False)
