"""Logical ranks vs processes.

The reference runs exactly one MPI rank per role (rank 0 = PS, ranks 1..P = workers; src/distributed_nn.py:87-133).
Here roles are *logical*: a job always has 1 PS + P workers, hosted by ``nprocs`` processes (one per GPU).  With
``nprocs == P + 1`` the mapping is the reference's; with fewer GPUs several workers (and the PS) share a process, which
is what lets the 1/2/4-GPU points of the benchmark run the same 1 PS + 7 workers job (strong scaling).
"""
from __future__ import annotations

from typing import Dict, List


class Placement:
    def __init__(self, num_workers: int, nprocs: int):
        self.num_workers, self.nprocs = num_workers, nprocs
        self.proc_of: Dict[int, int] = {0: 0}
        for w in range(1, num_workers + 1):
            self.proc_of[w] = w if nprocs >= num_workers + 1 else w % nprocs

    def local_workers(self, proc: int) -> List[int]:
        return [w for w in range(1, self.num_workers + 1) if self.proc_of[w] == proc]

    def worker_procs(self) -> List[int]:
        return sorted({self.proc_of[w] for w in range(1, self.num_workers + 1)})

    def active_procs(self) -> List[int]:
        return sorted(set(self.worker_procs()) | {0})

    def describe(self) -> str:
        parts = []
        for p in range(self.nprocs):
            roles = (["PS"] if p == 0 else []) + [f"w{w}" for w in self.local_workers(p)]
            parts.append(f"gpu{p}:[{','.join(roles) or 'idle'}]")
        return " ".join(parts)
