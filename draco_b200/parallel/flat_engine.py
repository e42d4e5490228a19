"""``--transport nccl_flat``: the HONEST library comparator for the fused transport.

`collective_engine.py` keeps the reference's message structure on purpose (62 broadcasts + 62 x P point-to-point messages per
ResNet-18 step, Python per-tensor decode) -- it measures what the reference's protocol costs on NCCL, not what a sensible
NCCL implementation costs.  This engine is the sensible one; it shares everything with the product EXCEPT the hand-written
communication / decode / model kernels, so that `bench.py` can show what those kernels buy once flat arenas and CUDA graphs
are factored out:

* flat parameter / gradient arenas (one element layout, parallel/arena.py);
* ONE ``dist.broadcast`` of the whole fp32 parameter arena per step, ONE flat message per logical worker upward
  (a single ``batch_isend_irecv``), nothing per tensor;
* decode = vectorised library ops over the whole arena: per-tensor exact-equality majority vote through a segmented
  mismatch count (no Python loop over tensors), mean for the plain mode; SGD-momentum as three flat ops;
* the workers' forward/backward + encode of ALL local workers is one CUDA graph (round-robin over the same concurrent
  worker streams the fused engine uses), replayed every step; NCCL calls stay outside the graph;
* model compute through the LIBRARY path (cuDNN convolutions, ATen BatchNorm / loss / Linear): `bench.py --impl nccl_flat`
  sets ``DRACO_CONV=cudnn DRACO_BN=aten DRACO_LINEAR=aten DRACO_FUSED_LOSS=0 DRACO_FUSED_PREP=0`` before building the job.

Supported: ``--approach baseline --mode normal`` and ``--approach maj_vote`` (the headline).  The cyclic code / geometric
median / Krum comparators stay on ``--transport nccl``.

Reference counterpart: the whole PS/worker step of src/master/rep_master.py:60-168 + src/worker/rep_worker.py.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.distributed as dist

from ..codes.adversary import generate_schedule
from ..config import JobConfig
from ..data import TensorDataset
from ..utils.metrics import limit_host_threads
from .arena import ArenaLayout
from .fused_engine import make_plan
from .placement import Placement
from .ps import build_codes
from .worker import WorkerCompute, make_model


class FlatNcclEngine:
    def __init__(self, cfg: JobConfig, rank: int, nprocs: int, device: torch.device, dataset: Optional[TensorDataset], group=None):
        if not (cfg.approach == "maj_vote" or (cfg.approach == "baseline" and cfg.mode == "normal")):
            raise ValueError("nccl_flat implements the repetition-code vote and the plain mean; use --transport nccl for "
                             f"approach={cfg.approach} mode={cfg.mode}")
        self.cfg, self.rank, self.nprocs, self.device, self.group = cfg, rank, nprocs, torch.device(device), group
        self.place = Placement(cfg.num_workers, nprocs)
        self.P = cfg.num_workers
        self.is_ps = rank == 0
        self.local_workers = self.place.local_workers(rank)
        self.groups, self.code = build_codes(cfg)
        self.step = 1
        limit_host_threads()
        if cfg.deterministic:
            torch.backends.cudnn.deterministic = True
            torch.backends.cudnn.benchmark = False
            os.environ.setdefault("CUBLAS_WORKSPACE_CONFIG", ":4096:8")
        model = make_model(cfg)
        self.layout = ArenaLayout.from_model(model, cfg.dtype == "bf16", channels_last=True)
        plan = make_plan(cfg, dataset, self.groups)
        self.worker = WorkerCompute(cfg, self.device, self.local_workers, plan, dataset, self.layout, None, model)
        self.params_f32 = self.worker.binder.params_f32
        D = self.layout.total
        self.sendbuf = {w: torch.zeros(D, dtype=torch.float32, device=self.device) for w in self.local_workers}
        self._send_views = {w: [self.layout.view(self.sendbuf[w], i) for i in range(self.layout.ntensors)] for w in self.local_workers}
        self.schedule = generate_schedule(self.P, cfg.worker_fail, cfg.max_steps)
        self.use_adv = cfg.err_mode != "none" and cfg.worker_fail > 0
        if self.use_adv and cfg.err_mode not in ("rev_grad", "constant"):
            raise ValueError("nccl_flat implements the rev_grad and constant attacks")
        # adversary bitmap per step on the device (read inside the captured graph: the step counter lives on the device too)
        bm = self.schedule.bitmaps().astype(np.int64)
        self.adv_bitmap = torch.from_numpy(bm).to(self.device) if self.use_adv else None
        self.step_dev = torch.ones(1, dtype=torch.int64, device=self.device)
        self.worker.step_dev = self.step_dev
        self.streams = [torch.cuda.Stream(self.device) for _ in range(min(max(int(cfg.worker_streams), 1), max(len(self.local_workers), 1)))] \
            if len(self.local_workers) > 1 and int(cfg.worker_streams) > 1 else []
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.kernels_per_step = 0
        if self.is_ps:
            self.slots = torch.zeros(self.P, D, dtype=torch.float32, device=self.device)
            self.momentum = torch.zeros(D, dtype=torch.float32, device=self.device)
            # element -> tensor id (segment id) for the per-tensor equality test
            seg = np.zeros(D, dtype=np.int64)
            for i, s in enumerate(self.layout.specs):
                seg[s.offset: s.offset + s.numel] = i
            self.seg = torch.from_numpy(seg).to(self.device)
            self.ntens = self.layout.ntensors
            # last arena element of every tensor's tile range (padding between tensors is zero in every slot: it never mismatches)
            nxt = [s.offset for s in self.layout.specs[1:]] + [D]
            self.seg_end = torch.tensor([n - 1 for n in nxt], dtype=torch.int64, device=self.device)

    # ------------------------------------------------------------------ worker side (captured)
    def _encode(self, w: int) -> None:
        buf = self.sendbuf[w]
        # gradients stay where autograd put them; ONE multi-tensor copy lays them out in the flat send buffer
        torch._foreach_copy_(self._send_views[w], self.worker.grad_refs[w][0])
        if self.use_adv:
            # liar(step, w) read from the device bitmap: buf = liar ? attack(buf) : buf  -- no host decision inside the graph
            bits = self.adv_bitmap[self.step_dev % self.adv_bitmap.numel()]        # same indexing as push_encode.cu
            liar = ((bits >> (w - 1)) & 1).to(torch.float32)
            mag = float(self.cfg.attack_magnitude)
            if self.cfg.err_mode == "rev_grad":
                buf.mul_(1.0 + liar * (mag - 1.0))
            else:
                buf.mul_(1.0 - liar).add_(liar * mag)

    def _enqueue_workers(self) -> None:
        from ..ops import norm as _norm
        if self.streams:
            main = torch.cuda.current_stream()
            fork = torch.cuda.Event()
            fork.record(main)
            try:
                for i, w in enumerate(self.local_workers):
                    st = self.streams[i % len(self.streams)]
                    _norm.UPDATE_RUNNING_STATS = (i == 0)
                    if i < len(self.streams):
                        st.wait_event(fork)
                    with torch.cuda.stream(st):
                        self.worker.forward_backward(w, None)
                        self._encode(w)
            finally:
                _norm.UPDATE_RUNNING_STATS = True
            for st in self.streams[: len(self.local_workers)]:
                main.wait_stream(st)
        else:
            for w in self.local_workers:
                self.worker.forward_backward(w, None)
                self._encode(w)

    # ------------------------------------------------------------------ PS side
    def _tensor_equal(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        """[ntensors] bool: tensors of rows a and b are bit-equal (NaN never equal, +0 == -0: torch.equal semantics)."""
        # segments are contiguous: mismatches per tensor = difference of an inclusive prefix sum at the segment ends
        c = torch.cumsum((a != b).to(torch.int32), 0, dtype=torch.int64)
        ends = c[self.seg_end]
        cnt = ends - torch.cat([ends.new_zeros(1), ends[:-1]])
        return cnt == 0

    def _decode(self) -> torch.Tensor:
        cfg = self.cfg
        if cfg.approach != "maj_vote" or cfg.mode == "normal":
            return self.slots.mean(0)
        agg = torch.zeros_like(self.slots[0])
        for members in self.groups.groups:                   # Boyer-Moore per tensor, vectorised over tensors
            rows = [m - 1 for m in members]
            cand = torch.zeros(self.ntens, dtype=torch.int64, device=self.device)
            count = torch.zeros(self.ntens, dtype=torch.int64, device=self.device)
            eq = {(i, j): self._tensor_equal(self.slots[rows[i]], self.slots[rows[j]]) for i in range(len(rows)) for j in range(i)}
            for k in range(len(rows)):
                same = torch.zeros(self.ntens, dtype=torch.bool, device=self.device)
                for c in range(k):
                    same |= (cand == c) & eq[(k, c)]
                zero = count == 0
                cand = torch.where(zero, torch.full_like(cand, k), cand)
                count = torch.where(zero, torch.ones_like(count), torch.where(same, count + 1, count - 1))
            sel = torch.tensor(rows, device=self.device)[cand][self.seg]            # winning row per element
            agg += self.slots.gather(0, sel.unsqueeze(0)).squeeze(0)
        return agg / len(self.groups.groups)

    def _apply(self, g: torch.Tensor) -> None:
        cfg = self.cfg
        if cfg.weight_decay:
            g = g.add(self.params_f32, alpha=cfg.weight_decay)
        if cfg.momentum:
            self.momentum.mul_(cfg.momentum).add_(g, alpha=1.0 - cfg.dampening)
            g = g.add(self.momentum, alpha=cfg.momentum) if cfg.nesterov else self.momentum
        self.params_f32.add_(g, alpha=-cfg.lr)

    # ------------------------------------------------------------------ the step
    def train_step(self, stage: bool = True) -> None:
        if self.nprocs > 1:
            dist.broadcast(self.params_f32, src=0, group=self.group)               # ONE flat broadcast
        if self.local_workers:
            if stage and self.worker.dataset is not None:
                self.worker.stage_batches(self.step)
            if self.graph is None and self.cfg.cuda_graphs and self.step >= 3:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.worker.binder.refresh_compute_copy()
                    self._enqueue_workers()
                self.graph = g
            if self.graph is not None:
                self.graph.replay()
            else:
                self.worker.binder.refresh_compute_copy()
                self._enqueue_workers()
        # ONE flat message per logical worker
        ops = []
        for w in range(1, self.P + 1):
            src = self.place.proc_of[w]
            if self.is_ps and src == 0:
                self.slots[w - 1].copy_(self.sendbuf[w])
            elif self.is_ps:
                ops.append(dist.P2POp(dist.irecv, self.slots[w - 1], src, group=self.group))
            elif src == self.rank:
                ops.append(dist.P2POp(dist.isend, self.sendbuf[w], 0, group=self.group))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        if self.is_ps:
            self._apply(self._decode())
        self.step_dev.add_(1)
        self.step += 1

    # ------------------------------------------------------------------ the engine protocol of Trainer / bench.py
    def read_metrics(self) -> Dict[str, float]:
        if not self.local_workers:
            return {}
        m = torch.stack([self.worker.metrics[w] for w in self.local_workers]).mean(0).tolist()
        return {"loss": m[0], "prec1": m[1], "prec5": m[2]}

    def enqueue_metrics_read(self):
        return self.read_metrics()

    def resolve_metrics(self, handle) -> Dict[str, float]:
        return handle

    def synchronize(self) -> None:
        torch.cuda.synchronize()

    def master_params(self) -> torch.Tensor:
        return self.params_f32

    def close(self) -> None:
        self.graph = None
