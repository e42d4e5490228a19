"""Repetition-code worker grouping.

Behavioural parity with the reference's ``group_assign`` / ``_assign`` /
``_group_identify`` (reference: src/util.py:69-97): workers ``1..P`` are cut into
consecutive groups of size ``r``; every member of a group trains on the same
batches (same data seed) so that honest members emit bit-identical gradients
and the PS can majority-vote (reference: src/master/rep_master.py:154-168).

Differences kept deliberately:
  * when ``P % r != 0`` the reference appends only rank ``P`` to the last
    group, silently dropping other remainder ranks; here *all* remainder
    ranks join the last group.
  * group seeds are drawn from a private ``RandomState`` seeded like the
    reference (428) instead of mutating numpy's global RNG.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np

SEED = 428  # reference: src/util.py:17


@dataclass(frozen=True)
class GroupPlan:
    """Static description of the repetition code for ``num_workers`` workers."""

    num_workers: int
    group_size: int
    groups: List[List[int]]            # groups[g] = sorted list of worker ranks (1-based)
    seeds: List[int]                   # seeds[g]  = data seed shared by group g
    rank_to_group: Dict[int, int] = field(default_factory=dict)

    @property
    def num_groups(self) -> int:
        return len(self.groups)

    def group_of(self, rank: int) -> int:
        """Group index of a worker rank; ``-1`` for the PS (rank 0)."""
        return -1 if rank == 0 else self.rank_to_group[rank]

    def member_index(self, rank: int) -> int:
        return self.groups[self.rank_to_group[rank]].index(rank)

    def tolerance(self, g: int) -> int:
        """Number of liars group ``g`` can out-vote: floor((r_g - 1) / 2)."""
        return (len(self.groups[g]) - 1) // 2

    def as_table(self) -> np.ndarray:
        """``[num_groups, max_group]`` int32 table of 0-based worker slots, -1 padded.

        This is the layout the vote kernel consumes (csrc/cuda/vote.cu).
        """
        width = max(len(g) for g in self.groups)
        tab = -np.ones((self.num_groups, width), dtype=np.int32)
        for gi, g in enumerate(self.groups):
            tab[gi, : len(g)] = np.asarray(g, dtype=np.int32) - 1
        return tab


def group_assign(num_workers: int, group_size: int, seed: int = SEED) -> GroupPlan:
    """Build the repetition-code plan for ``P = num_workers`` and ``r = group_size``."""
    if num_workers < 1:
        raise ValueError("need at least one worker")
    if group_size < 1:
        raise ValueError("group size must be >= 1")
    group_size = min(group_size, num_workers)
    k = num_workers // group_size
    groups = [list(range(g * group_size + 1, (g + 1) * group_size + 1)) for g in range(k)]
    for r in range(k * group_size + 1, num_workers + 1):
        groups[-1].append(r)
    rng = np.random.RandomState(seed)
    seeds = [int(rng.randint(0, 20000)) for _ in groups]
    rank_to_group = {r: gi for gi, g in enumerate(groups) for r in g}
    return GroupPlan(num_workers, group_size, groups, seeds, rank_to_group)
