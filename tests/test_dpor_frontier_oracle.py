"""The frontier ("wide") DPOR restatement (oracle/dpor_frontier.c) against the sequential DPORwHeuristics restatement
(oracle/dpor.c): width 1 is the reference's order; at any width a bounded-exhaustive single-rank search visits the
same set of schedules; several ranks (rank-local explored sets + the deterministic steal protocol) run more
interleavings and cover most of that set."""
import numpy as np

import demi_b200 as D
from demi_b200 import _native as N


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


def test_width_one_is_the_sequential_order(oracle):
    rng = np.random.default_rng(1)
    for prog in raft_programs(10, rng):
        ext = D.pack_externals(prog)
        for maxm, maxi, flags in [(24, 400, 1), (100, 120, 3)]:
            rc, r, cviol, ch = oracle.dpor_search(N.MODEL_RAFT5, ext, maxm, maxi, 0, 0, -1, model_flags=flags, node_cap=1 << 15,
                                                  explored_slots=1 << 20, heap_cap=1 << 20)
            F = oracle.frontier_params(maxm, maxi, 1, explored_slots=1 << 20, pool_cap=1 << 21)
            rc2, res, vs, hs = oracle.dpor_frontier(N.MODEL_RAFT5, ext, F, 1, model_flags=flags)
            assert rc == 0 and rc2 == 0
            assert len(ch) == len(hs[0]) and (ch == hs[0]).all()                # schedule by schedule
            assert (cviol["schedule_hash"] == vs[0]["schedule_hash"]).all()
            assert res[0]["exhausted"] == r["exhausted"] and res[0]["budget_exhausted"] == r["budget_exhausted"]


def exhaustive_cases():
    rng = np.random.default_rng(11)
    cases = [(N.MODEL_RAFT5, prog, 36, 3) for prog in raft_programs(4, rng)]
    for n in (6, 8):
        prog = [D.Start(a) for a in range(3)] + [D.Send(k % 3, 1, k) for k in range(n)]
        cases.append((N.MODEL_PINGPONG3, prog, 2 * n + 2, 1 | (2 << 8)))
    return cases


def test_exhaustive_set_parity_and_multi_rank_coverage(oracle):
    with_viol = 0
    for model, prog, maxm, flags in exhaustive_cases():
        ext = D.pack_externals(prog)
        rc, r, cviol, ch = oracle.dpor_search(model, ext, maxm, 100000, 0, 0, -1, model_flags=flags, node_cap=1 << 16,
                                              explored_slots=1 << 22, heap_cap=1 << 22, cap_viol=100000)
        assert r["exhausted"] == 1
        seq_all = set(int(x) for x in ch)
        seq_viol = set(int(x) for x in cviol["schedule_hash"])
        with_viol += bool(seq_viol)
        for width, ranks in [(8, 1), (256, 1), (64, 2), (16, 4)]:
            F = oracle.frontier_params(maxm, 100000, width, explored_slots=1 << 22, pool_cap=1 << 22, rounds_per_exchange=2,
                                       steal_max=256)
            rc2, res, vs, hs = oracle.dpor_frontier(model, ext, F, ranks, model_flags=flags, cap_viol=100000)
            assert rc2 == 0 and (res["exhausted"] == 1).all()
            f_all = set(int(x) for h in hs for x in h)
            f_viol = set(int(x) for v in vs for x in v["schedule_hash"])
            if ranks == 1:
                assert f_all == seq_all and f_viol == seq_viol
            else:
                # every rank keeps its own explored set and learns the other ranks' newly explored pairs at the exchange
                # points: every racing pair is still reversed about once (the work stays close to the sequential search's),
                # but WHICH context reverses it depends on where a point ran, so the visited schedules differ
                assert len(f_viol & seq_viol) >= 0.8 * len(seq_viol)
                assert len(seq_all) * 0.8 <= res["interleavings"].sum() <= 3 * len(seq_all)
                assert res["records_sent"].sum() == res["records_received"].sum() > 0
    assert with_viol >= 2


def test_multi_rank_runs_are_deterministic_and_respect_the_budget(oracle):
    prog = D.raft5_program(client_cmds=2)[:-1]
    ext = D.pack_externals(prog)
    for flags, budget in [(0, 120), (N.FR_NO_HISTORY, 3000)]:               # trackHistory = true / false
        F = oracle.frontier_params(60, budget, 32, explored_slots=1 << 20, pool_cap=1 << 22, rounds_per_exchange=2, steal_max=64,
                                   flags=flags)
        a = oracle.dpor_frontier(N.MODEL_RAFT5, ext, F, 4, model_flags=3)
        b = oracle.dpor_frontier(N.MODEL_RAFT5, ext, F, 4, model_flags=3)
        assert a[0] == 0 and (a[1] == b[1]).all()
        assert all((x == y).all() for x, y in zip(a[3], b[3]))
        assert a[1]["interleavings"].sum() == budget and (a[1]["budget_exhausted"] == 1).all()
        assert (a[1]["interleavings"] > 0).sum() >= (4 if flags else 2)         # the work is spread (a small budget goes to the first ranks)
