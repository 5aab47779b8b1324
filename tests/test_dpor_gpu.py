"""GPU parity for K3: batched DPORwHeuristics searches vs the sequential CPU oracle, bit-exact per search."""
import numpy as np
import pytest

import demi_b200 as D
from demi_b200 import _native as N

pytestmark = pytest.mark.gpu

FIELDS = ["interleavings", "violations", "deliveries", "races", "n_nodes", "n_explored", "heap_left", "exhausted",
          "budget_exhausted", "status"]


def raft_programs(n, rng):
    """Independent searches: different boot orders / client commands (one DPOR per external subsequence)."""
    progs = []
    for _ in range(n):
        order = rng.permutation(5)
        ev = [D.Start(int(a)) for a in order]
        k = int(rng.integers(3, 6))
        ev += [D.Send(int(a), 1, 0x1F) for a in rng.permutation(5)[:k]]
        ev += [D.Send(int(rng.integers(0, 5)), 2, int(rng.integers(1, 50))) for _ in range(int(rng.integers(0, 3)))]
        progs.append(ev)
    return progs


def check(model, progs, flags, maxm, maxi, oracle, looking_for=0, stop=False, depth_bound=-1, **caps):
    eng = D.Engine(D.SchedulerConfig(model, model_flags=flags))
    res, viol, hashes = eng.dpor_batch(progs, maxm, maxi, looking_for, stop, depth_bound, cap_viol=32, want_hashes=True, **caps)
    for s, prog in enumerate(progs):
        rc, r, cviol, chashes = oracle.dpor_search(model, D.pack_externals(prog), maxm, maxi, looking_for, 1 if stop else 0,
                                                    depth_bound, model_flags=flags,
                                                    node_cap=caps.get("node_cap", 4096),
                                                    explored_slots=caps.get("explored_slots", 1 << 16),
                                                    heap_cap=caps.get("heap_cap", 1 << 15), cap_viol=32)
        for f in FIELDS:
            assert res[s][f] == r[f], (s, f, res[s], r)
        n = int(r["interleavings"])
        assert (hashes[s][:n] == chashes).all(), s           # the same schedules in the same order
        nv = min(int(r["violations"]), 32)
        assert (viol[s][:nv] == cviol[:nv]).all()
    return res, viol, hashes


def test_dpor_raft_exhaustive_small_bound(oracle):
    rng = np.random.default_rng(1)
    res, _, hashes = check(N.MODEL_RAFT5, raft_programs(40, rng), 1, 24, 400, oracle)
    assert (res["status"] == 0).all() and res["exhausted"].sum() > 0
    assert len(set(int(x) for x in res["interleavings"])) > 3          # searches of different sizes


def test_dpor_raft_depth100_budgeted(oracle):
    rng = np.random.default_rng(2)
    res, _, _ = check(N.MODEL_RAFT5, raft_programs(12, rng), 3, 100, 120, oracle, heap_cap=1 << 17)
    assert (res["budget_exhausted"] == 1).all() and (res["status"] == 0).all()


def test_dpor_finds_pingpong_violation_and_stops(oracle):
    flags = 1 | (2 << 8)     # violation 7 once actor 0 has received 2 pongs
    progs = []
    rng = np.random.default_rng(3)
    for _ in range(20):
        ev = [D.Start(a) for a in range(3)]
        ev += [D.Send(int(rng.integers(0, 3)), 1, k) for k in range(int(rng.integers(4, 9)))]
        progs.append(ev)
    res, viol, _ = check(N.MODEL_PINGPONG3, progs, flags, 60, 500, oracle, looking_for=7, stop=True)
    assert res["violations"].sum() > 0
    # without stopIfViolationFound the search runs on and collects the whole violating set
    res2, viol2, _ = check(N.MODEL_PINGPONG3, progs, flags, 60, 500, oracle, looking_for=7, stop=False)
    assert (res2["violations"] >= res["violations"]).all() and res2["interleavings"].sum() > res["interleavings"].sum()


def test_dpor_depth_bound_and_capacity_status(oracle):
    rng = np.random.default_rng(4)
    check(N.MODEL_RAFT5, raft_programs(6, rng), 1, 40, 200, oracle, depth_bound=4)
    # a tiny heap overflows loudly and identically
    res, _, _ = check(N.MODEL_RAFT5, raft_programs(4, rng), 1, 100, 200, oracle, heap_cap=64)
    assert (res["status"] == 4).any()


def test_dpor_scheduler_mirror(oracle):
    cfg = D.SchedulerConfig(N.MODEL_PINGPONG3, model_flags=1 | (2 << 8))
    dpor = D.DPORwHeuristics(cfg, stopIfViolationFound=True, max_interleavings=300)
    with pytest.raises(ValueError):
        dpor.test([D.Start(0)], 7)
    dpor.setMaxMessagesToSchedule(40)
    prog = [D.Start(a) for a in range(3)] + [D.Send(2, 1, k) for k in range(3)] + [D.WaitQuiescence()]
    hit = dpor.test(prog, 7)
    assert hit is not None and hit["code"] == 7
    assert dpor.test(prog, 5) is None
    eng = D.Engine(cfg)
    with pytest.raises(D.DemiError):                      # "unsuported external event" (DPORwHeuristics.scala:710)
        eng.dpor_batch([[D.Start(0), D.Kill(0)]], 10, 10)
