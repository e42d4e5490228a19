"""Property-based tests (hypothesis) of the host-side invariants the device kernels rely on."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from draco_b200 import _native as N
from draco_b200.codes import adversary, cyclic, oracle
from draco_b200.codes.repetition import group_assign
from draco_b200.data import BatchPlan
from draco_b200.parallel.placement import Placement
from draco_b200.utils import codec

COMMON = dict(deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])


@settings(max_examples=40, **COMMON)
@given(P=st.integers(1, 24), r=st.integers(1, 9))
def test_group_assign_is_a_partition(P, r):
    r = min(r, P)
    plan = group_assign(P, r)
    flat = [w for g in plan.groups for w in g]
    assert sorted(flat) == list(range(1, P + 1))                     # every worker in exactly one group, PS in none
    assert all(len(g) >= r for g in plan.groups) and len(plan.groups) == max(P // r, 1)
    assert len(plan.seeds) == len(plan.groups)
    for gi, g in enumerate(plan.groups):
        for mi, w in enumerate(g):
            assert plan.group_of(w) == gi and plan.member_index(w) == mi


@settings(max_examples=40, **COMMON)
@given(P=st.integers(1, 16), nprocs=st.integers(1, 17))
def test_placement_hosts_every_logical_rank_once(P, nprocs):
    pl = Placement(P, nprocs)
    hosted = sorted(w for p in range(nprocs) for w in pl.local_workers(p))
    assert hosted == list(range(1, P + 1)) and pl.proc_of[0] == 0
    assert set(pl.active_procs()) == set(pl.worker_procs()) | {0}
    if nprocs >= P + 1:                                              # the reference's one-rank-per-role layout
        assert all(pl.proc_of[w] == w for w in range(1, P + 1))


@settings(max_examples=25, **COMMON)
@given(P=st.integers(1, 12), s=st.integers(0, 5), steps=st.integers(1, 40))
def test_adversary_schedule_draws_distinct_workers(P, s, steps):
    s = min(s, P)
    sched = adversary.generate_schedule(P, s, steps)
    bm = sched.bitmaps()
    for t in range(1, steps + 1):
        liars = [w for w in range(1, P + 1) if sched.is_adversary(w, t)]
        assert len(liars) == s
        assert int(bm[t]) == sum(1 << (w - 1) for w in liars)       # the device bitmap says the same thing


@settings(max_examples=25, **COMMON)
@given(data=st.data())
def test_cyclic_code_recovers_the_sum_for_random_attacks(data):
    n = data.draw(st.integers(3, 9))
    s = data.draw(st.integers(1, (n - 1) // 2))
    k = data.draw(st.integers(0, s))
    liars = data.draw(st.lists(st.integers(0, n - 1), min_size=k, max_size=k, unique=True))
    seed = data.draw(st.integers(0, 2 ** 16))
    rng = np.random.RandomState(seed)
    d = 16
    c = cyclic.search_w(n, s)
    G = rng.randn(n, d)
    f = rng.randn(d) + 1.0
    R = np.stack([cyclic.encode(c, i, G) for i in range(n)])
    for l in liars:
        R[l] += (10.0 + 100.0 * rng.rand()) * rng.choice([-1.0, 1.0]) * (1.0 + rng.rand(d))
    dec, healthy = cyclic.decode(c, R, f)
    assert not set(healthy) & set(liars)
    assert np.abs(dec - G.sum(0)).max() < 1e-6 * max(1.0, np.abs(G.sum(0)).max())
    # the C++ locator (the code the PS kernel shares) agrees on who is healthy
    E = (R @ f).astype(np.complex128)
    Eh = np.stack([E.real, E.imag], axis=1).copy()
    v = np.zeros((n, 2)); mask = np.zeros(1, dtype=np.uint32); fl = np.zeros(1, dtype=np.int32)
    N.check(N.host().drc_host_locate(Eh.ctypes.data, 1, n, s, 1e-4, v.ctypes.data, mask.ctypes.data, fl.ctypes.data), "locate")
    flagged = {i for i in range(n) if not (int(mask[0]) >> i) & 1}
    assert set(liars) <= flagged and len(flagged) <= 2 * s
    vt = v[:, 0] + 1j * v[:, 1]
    assert np.abs((vt @ R).real - G.sum(0)).max() < 1e-6 * max(1.0, np.abs(G.sum(0)).max())


@settings(max_examples=30, **COMMON)
@given(data=st.data())
def test_majority_vote_returns_the_honest_value_when_liars_are_a_minority(data):
    r = data.draw(st.integers(1, 9))
    k = data.draw(st.integers(0, (r - 1) // 2))
    liars = set(data.draw(st.lists(st.integers(0, r - 1), min_size=k, max_size=k, unique=True)))
    rng = np.random.RandomState(data.draw(st.integers(0, 999)))
    honest = rng.randn(12).astype(np.float32)
    members = [honest.copy() if i not in liars else (honest * -100 + rng.randn(12)).astype(np.float32) for i in range(r)]
    w = oracle.boyer_moore_winner(members)
    assert np.array_equal(members[w], honest)


@settings(max_examples=30, **COMMON)
@given(data=st.data())
def test_codec_roundtrips_any_tensor(data):
    dtype = data.draw(st.sampled_from([np.float32, np.float64, np.complex64, np.int32, np.uint8]))
    n = data.draw(st.integers(0, 3000))
    kind = data.draw(st.sampled_from(["gauss", "zeros", "const", "sparse", "bytes"]))
    rng = np.random.RandomState(data.draw(st.integers(0, 9999)))
    if kind == "gauss":
        a = (rng.randn(n) * 10.0 ** rng.randint(-6, 6))
    elif kind == "zeros":
        a = np.zeros(n)
    elif kind == "const":
        a = np.full(n, -100.0)
    elif kind == "sparse":
        a = rng.randn(n) * (rng.rand(n) < 0.05)
    else:
        a = rng.randint(0, 256, size=n)
    a = a.astype(dtype)
    blob = codec.compress(a)
    b = codec.decompress(blob)
    assert b.dtype == a.dtype and b.shape == a.shape and a.tobytes() == b.tobytes()     # lossless, bit for bit


@settings(max_examples=25, **COMMON)
@given(approach=st.sampled_from(["baseline", "maj_vote", "cyclic"]), step=st.integers(1, 500), B=st.sampled_from([4, 32, 128]),
       P=st.integers(3, 9))
def test_batch_plan_gives_holders_of_a_batch_the_same_indices(approach, step, B, P):
    size = 2048
    kw = {}
    if approach == "maj_vote":
        plan = group_assign(P, 3)
        kw = dict(group_of={w: plan.group_of(w) for w in range(1, P + 1)}, group_seeds=plan.seeds)
    red = 3 if approach == "cyclic" else 1
    bp = BatchPlan(approach=approach, dataset_size=size, batch_size=B, num_workers=P, redundancy=red, **kw)
    seen = {}
    for w in range(1, P + 1):
        ids = bp.batch_ids(step, w)
        idxs = bp.indices(step, w)
        assert len(ids) == len(idxs) == red
        for bid, idx in zip(ids, idxs):
            idx = np.asarray(idx)
            assert idx.shape == (B,) and idx.min() >= 0 and idx.max() < size
            if bid in seen:
                assert np.array_equal(seen[bid], idx)                  # same batch id => same samples, whoever holds it
            seen[bid] = idx
    if approach == "maj_vote":
        g = group_assign(P, 3)
        for grp in g.groups:
            assert len({bp.batch_ids(step, w)[0] for w in grp}) == 1   # a group trains on one batch
    if approach == "cyclic":
        assert len(seen) == P                                          # n batches, each held by 2s+1 workers


@settings(max_examples=20, **COMMON)
@given(P=st.integers(2, 9), d=st.integers(1, 200), seed=st.integers(0, 999))
def test_weight_space_geometric_median_matches_weiszfeld(P, d, seed):
    rng = np.random.RandomState(seed)
    X = rng.randn(P, d) * rng.choice([1e-3, 1.0, 50.0])
    if P > 2:
        X[rng.randint(P)] *= -100.0
    w = oracle.geometric_median_weights(X, eps=1e-13, max_iter=4000)
    ref = oracle.geometric_median(X, eps=1e-13, max_iter=20000)
    assert abs(w.sum() - 1.0) < 1e-9 and w.min() >= 0
    scale = max(1.0, np.abs(X).max())
    assert np.abs(w @ X - ref).max() < 1e-5 * scale
