"""GPU parity for K3F: one DPORwHeuristics search explored as a frontier of backtrack points (demi_dpor_frontier)
against its CPU restatement (oracle/dpor_frontier.c), and set parity with the sequential reference order."""
import numpy as np
import pytest

import demi_b200 as D
from demi_b200 import _native as N

pytestmark = pytest.mark.gpu

FIELDS = ["interleavings", "violations", "deliveries", "races", "keys_enqueued", "keys_dropped", "explored_pairs",
          "pool_left", "rounds", "exhausted", "budget_exhausted", "status", "trace_slots"]


def raft_programs(n, rng):
    progs = []
    for _ in range(n):
        order = rng.permutation(5)
        ev = [D.Start(int(a)) for a in order]
        k = int(rng.integers(3, 6))
        ev += [D.Send(int(a), 1, 0x1F) for a in rng.permutation(5)[:k]]
        ev += [D.Send(int(rng.integers(0, 5)), 2, int(rng.integers(1, 50))) for _ in range(int(rng.integers(0, 3)))]
        progs.append(ev)
    return progs


def check(model, prog, flags, maxm, maxi, width, oracle, **kw):
    if "fr_flags" in kw:
        kw["flags"] = kw.pop("fr_flags")
    eng = D.Engine(D.SchedulerConfig(model, model_flags=flags))
    F = eng.frontier_params(maxm, maxi, width, **kw)
    r, viol, hashes = eng.dpor_frontier(prog, F)
    OF = oracle.frontier_params(maxm, maxi, width, **{k: v for k, v in kw.items() if k in ("explored_slots", "pool_cap", "stop_if_found", "looking_for", "flags")})
    rc, ores, oviol, ohashes = oracle.dpor_frontier(model, D.pack_externals(prog), OF, 1, model_flags=flags)
    assert rc == 0
    for f in FIELDS:
        assert r[f] == ores[0][f], (f, r, ores[0])
    assert (hashes == ohashes[0]).all()                         # the same schedules in the same (slot) order
    assert (viol == oviol[0]).all()
    return r, viol, hashes


@pytest.mark.parametrize("width", [1, 7, 64, 1024])
def test_frontier_matches_oracle_raft_budgeted(oracle, width):
    rng = np.random.default_rng(3)
    for prog in raft_programs(3, rng):
        r, _, _ = check(N.MODEL_RAFT5, prog, 3, 60, 150, width, oracle, explored_slots=1 << 20, pool_cap=1 << 21)
        assert r["status"] == 0 and r["budget_exhausted"] == 1
        r, _, _ = check(N.MODEL_RAFT5, prog, 3, 100, 5000, width, oracle, explored_slots=1 << 20, pool_cap=1 << 22)
        assert r["status"] == 0 and r["exhausted"] == 1 and r["interleavings"] > 300


@pytest.mark.parametrize("width", [1, 16, 512])
def test_frontier_without_history_matches_oracle(oracle, width):
    """trackHistory = false (DPORwHeuristics.scala:86): every backtrack point is replayed; only the budget ends it."""
    prog = D.raft5_program(client_cmds=2)[:-1]
    r, _, hashes = check(N.MODEL_RAFT5, prog, 3, 50, 1200, width, oracle, pool_cap=1 << 22, fr_flags=N.FR_NO_HISTORY)
    assert r["budget_exhausted"] == 1 and r["explored_pairs"] == 0 and r["keys_dropped"] == 0


def test_frontier_width1_is_the_reference_order(oracle):
    """width = 1 dequeues one backtrack point per race scan: the sequence of schedules is the sequential
    DPORwHeuristics restatement's (oracle/dpor.c), interleaving by interleaving."""
    rng = np.random.default_rng(4)
    for prog in raft_programs(4, rng):
        eng = D.Engine(D.SchedulerConfig(N.MODEL_RAFT5, model_flags=1))
        F = eng.frontier_params(24, 400, 1, explored_slots=1 << 18, pool_cap=1 << 19)
        r, viol, hashes = eng.dpor_frontier(prog, F)
        rc, sr, sviol, shashes = oracle.dpor_search(N.MODEL_RAFT5, D.pack_externals(prog), 24, 400, 0, 0, -1, model_flags=1,
                                                    node_cap=1 << 14, explored_slots=1 << 18, heap_cap=1 << 18)
        assert len(hashes) == len(shashes) and (hashes == shashes).all()
        assert r["exhausted"] == sr["exhausted"]


def test_frontier_exhaustive_set_parity_with_the_sequential_search(oracle):
    """Bounded-exhaustive searches (stopIfViolationFound = false, fixed max_messages): the set of schedules — and so
    the set of violating schedules — a wide frontier visits equals the sequential search's (oracle_dpor_search)."""
    cases = []
    rng = np.random.default_rng(5)
    for prog in raft_programs(3, rng):
        cases.append((N.MODEL_RAFT5, prog, 3, 40))
    for n in (6, 8):
        prog = [D.Start(a) for a in range(3)] + [D.Send(k % 3, 1, k) for k in range(n)]
        cases.append((N.MODEL_PINGPONG3, prog, 1 | (2 << 8), 2 * n + 2))
    n_viol = 0
    for model, prog, flags, maxm in cases:
        rc, sr, sviol, shashes = oracle.dpor_search(model, D.pack_externals(prog), maxm, 100000, 0, 0, -1, model_flags=flags,
                                                    node_cap=1 << 16, explored_slots=1 << 22, heap_cap=1 << 22, cap_viol=100000)
        assert sr["exhausted"] == 1
        for width in (16, 4096):
            eng = D.Engine(D.SchedulerConfig(model, model_flags=flags))
            F = eng.frontier_params(maxm, 100000, width, explored_slots=1 << 22, pool_cap=1 << 22)
            r, viol, hashes = eng.dpor_frontier(prog, F, cap_viol=100000)
            assert r["exhausted"] == 1 and r["status"] == 0
            assert set(int(x) for x in hashes) == set(int(x) for x in shashes)
            assert set(int(x) for x in viol["schedule_hash"]) == set(int(x) for x in sviol["schedule_hash"])
        n_viol += len(sviol)
    assert n_viol > 0


def test_frontier_stop_if_found_and_bcast(oracle):
    flags = 1 | (2 << 8)
    prog = [D.Start(a) for a in range(3)] + [D.Send(k % 3, 1, k) for k in range(6)]
    r, viol, _ = check(N.MODEL_PINGPONG3, prog, flags, 14, 10000, 8, oracle, stop_if_found=True)
    assert r["violations"] >= 1
    check(N.MODEL_BCAST32, D.bcast32_program(2)[:-1], 0, 40, 300, 32, oracle, explored_slots=1 << 20, pool_cap=1 << 21)


def test_frontier_capacity_status(oracle):
    eng = D.Engine(D.SchedulerConfig(N.MODEL_RAFT5, model_flags=1))
    prog = D.raft5_program()[:-1]
    F = eng.frontier_params(60, 5000, 64, explored_slots=1 << 20, pool_cap=256)
    with pytest.raises(D.DemiError):
        eng.dpor_frontier(prog, F)
    assert eng.last_frontier["status"] == 4          # DEMI_DS_HEAP_OVF
