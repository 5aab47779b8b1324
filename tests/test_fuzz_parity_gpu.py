"""GPU parity: the CUDA RandomScheduler engine vs the CPU oracle, bit-exact, through the C ABI."""
import numpy as np
import pytest

import demi_b200 as D
from demi_b200 import _native as N

pytestmark = pytest.mark.gpu

FIELDS = ["violation", "steps", "state_hash", "trace_hash", "n_nodes", "n_events", "max_pending", "status"]


def assert_same(gpu, cpu):
    for f in FIELDS:
        bad = np.nonzero(gpu[f] != cpu[f])[0]
        assert len(bad) == 0, "field %s differs at %d prefixes, first idx %d: gpu=%s cpu=%s" % (
            f, len(bad), bad[0], gpu[bad[0]], cpu[bad[0]])


CASES = [
    # model, program, model_flags, max_messages, interval, n, blocked_mask, ignore_timers
    ("raft5-bug-d50", N.MODEL_RAFT5, lambda: D.raft5_program(), 1, 50, 5, 20000, 0, 0),
    ("raft5-nobug-d50", N.MODEL_RAFT5, lambda: D.raft5_program(), 0, 50, 5, 5000, 0, 0),
    ("raft5-bugs-clients-d50", N.MODEL_RAFT5, lambda: D.raft5_program(client_cmds=3), 3, 50, 5, 5000, 0, 0),
    ("raft5-d100-int30", N.MODEL_RAFT5, lambda: D.raft5_program(), 1, 100, 30, 3000, 0, 0),
    ("raft5-unbounded-notimers", N.MODEL_RAFT5, lambda: D.raft5_program(client_cmds=2), 1, -1, 0, 500, 0, 1),
    ("raft5-blocked", N.MODEL_RAFT5, lambda: D.raft5_program(), 1, 50, 5, 3000, 0b00100, 0),
    ("pingpong3-c1", N.MODEL_PINGPONG3, lambda: D.pingpong3_program(100), 0, -1, 0, 2000, 0, 0),
    ("pingpong3-viol", N.MODEL_PINGPONG3, lambda: D.pingpong3_program(100), 1 | (10 << 8), -1, 7, 500, 0, 0),
    ("pingpong3-blocked", N.MODEL_PINGPONG3, lambda: D.pingpong3_program(40), 0, -1, 0, 500, 0b010, 0),
    ("bcast32-ttl2", N.MODEL_BCAST32, lambda: D.bcast32_program(2), 0, 200, 0, 300, 0, 0),
    ("bcast32-viol", N.MODEL_BCAST32, lambda: D.bcast32_program(3), 4, 200, 10, 300, 0, 0),
]


@pytest.fixture(params=["lane+warp", "warp-only"])
def engine_mode(request, monkeypatch):
    """Both engines must be bit-exact: K1-lane (+ deferred prefixes on the warp engine) and the warp engine alone."""
    if request.param == "warp-only":
        monkeypatch.setenv("DEMI_DISABLE_LANE_ENGINE", "1")
    else:
        monkeypatch.delenv("DEMI_DISABLE_LANE_ENGINE", raising=False)
    return request.param


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_fuzz_batch_matches_oracle(case, oracle, engine_mode):
    _, model, prog, flags, maxm, interval, n, blocked, ignore = case
    ext = D.pack_externals(prog())
    eng = D.Engine(D.SchedulerConfig(model, model_flags=flags, blocked_mask=blocked, ignoreTimers=bool(ignore)))
    eng.set_externals(ext)
    gpu = eng.fuzz_batch(1, n, maxm, interval)
    cpu = oracle.fuzz_batch(model, ext, 1, n, maxm, interval, model_flags=flags, blocked_mask=blocked,
                            ignore_timers=ignore)
    assert_same(gpu, cpu)
    assert (gpu["status"] == 0).all()


def test_partition_kill_segments(oracle, engine_mode):
    """Kill / Partition / UnPartition between quiescent segments (EventOrchestrator.scala:132-189)."""
    ev = [D.Start(a) for a in range(3)]
    ev += [D.Send(k % 3, 1, k) for k in range(12)]
    ev += [D.WaitQuiescence(), D.Partition(0, 1), D.Kill(2)]
    ev += [D.Send(k % 3, 1, 100 + k) for k in range(12)]
    ev += [D.WaitQuiescence(), D.UnPartition(1, 0), D.UnPartition(0, 1), D.Start(2)]
    ev += [D.Send(k % 3, 1, 200 + k) for k in range(9)]
    ev += [D.WaitQuiescence()]
    ext = D.pack_externals(ev)
    eng = D.Engine(D.SchedulerConfig(N.MODEL_PINGPONG3))
    eng.set_externals(ext)
    gpu = eng.fuzz_batch(7, 1000, -1, 0)
    cpu = oracle.fuzz_batch(N.MODEL_PINGPONG3, ext, 7, 1000, -1, 0)
    assert_same(gpu, cpu)


@pytest.mark.parametrize("model,prog,flags,maxm,interval", [
    (N.MODEL_RAFT5, lambda: D.raft5_program(client_cmds=2), 1, 50, 5),
    (N.MODEL_PINGPONG3, lambda: D.pingpong3_program(30), 0, -1, 0),
    (N.MODEL_BCAST32, lambda: D.bcast32_program(2), 0, 100, 0),
])
def test_event_trace_and_dep_tree_match_oracle(model, prog, flags, maxm, interval, oracle):
    """Full EventTrace (EventTrace.scala:20) and DepTracker tree of single executions."""
    ext = D.pack_externals(prog())
    eng = D.Engine(D.SchedulerConfig(model, model_flags=flags))
    eng.set_externals(ext)
    for seed in (1, 2, 3, 12345, -7):
        gev, gpar, gres = eng.fuzz_trace(seed, maxm, interval)
        cev, cpar, cres = oracle.fuzz_trace(model, ext, seed, maxm, interval, model_flags=flags)
        assert len(gev) == len(cev) and len(gpar) == len(cpar)
        assert (gev == cev).all()
        assert (gpar == cpar).all()
        for f in FIELDS:
            assert gres[f] == cres[f], f
        batch = eng.fuzz_batch(seed, 1, maxm, interval)[0]
        assert batch["trace_hash"] == gres["trace_hash"]


def test_random_scheduler_explore_finds_first_violation(oracle):
    cfg = D.SchedulerConfig(N.MODEL_RAFT5, model_flags=1)
    sched = D.RandomScheduler(cfg, max_executions=4000, invariant_check_interval=5, seed=1)
    sched.setMaxMessages(50)
    prog = D.raft5_program()
    found = sched.explore(prog)
    cpu = oracle.fuzz_batch(N.MODEL_RAFT5, D.pack_externals(prog), 1, 4000, 50, 5, model_flags=1)
    hits = np.nonzero(cpu["violation"])[0]
    assert len(hits) > 0 and found is not None
    trace, code = found
    assert code == cpu["violation"][hits[0]]
    cev, _, _ = oracle.fuzz_trace(N.MODEL_RAFT5, D.pack_externals(prog), 1 + int(hits[0]), 50, 5, model_flags=1)
    assert (trace == cev).all()
    # looking for a code that does not occur => None (violationMatches, RandomScheduler.scala:138-154)
    assert sched.test(prog, 99) is None


def test_large_batch_properties():
    """BASELINE config[1] sized batch: size-independent properties (no oracle at this size)."""
    eng = D.Engine(D.SchedulerConfig(N.MODEL_RAFT5, model_flags=1))
    eng.set_externals(D.raft5_program())
    n = 2_000_000
    a = eng.fuzz_batch(1, n, 50, 5)
    st = eng.stats()
    assert st.deliveries == int(a["steps"].sum()) and st.violations == int((a["violation"] != 0).sum())
    assert (a["status"] == 0).all()
    assert (a["steps"] <= 51).all() and (a["steps"] >= 1).all()
    assert ((a["violation"] == 0) | (a["steps"] % 5 == 0)).all()       # violations only at check points
    assert ((a["violation"] != 0) | (a["steps"] == 51)).all()           # otherwise runs to maxMessages+1
    # determinism + seed-offset consistency: prefix i of base s == prefix 0 of base s+i
    b = eng.fuzz_batch(1 + 1_000_000, 1000, 50, 5)
    assert (a[1_000_000:1_001_000] == b).all()
    # correct Raft (no seeded bug) never violates
    eng2 = D.Engine(D.SchedulerConfig(N.MODEL_RAFT5, model_flags=0))
    eng2.set_externals(D.raft5_program())
    c = eng2.fuzz_batch(1, 500_000, 50, 5)
    assert (c["violation"] == 0).all() and (c["steps"] == 51).all()


FIFO_CASES = [
    ("raft5", N.MODEL_RAFT5, lambda: D.raft5_program(client_cmds=2), 1, 50, 5, 4000, 0),
    ("raft5-blocked", N.MODEL_RAFT5, lambda: D.raft5_program(), 1, 50, 5, 1500, 0b01000),
    ("raft5-d100", N.MODEL_RAFT5, lambda: D.raft5_program(), 3, 100, 30, 1000, 0),
    ("pingpong3", N.MODEL_PINGPONG3, lambda: D.pingpong3_program(60), 0, -1, 0, 1000, 0),
    ("pingpong3-blocked", N.MODEL_PINGPONG3, lambda: D.pingpong3_program(30), 0, -1, 0, 500, 0b100),
    ("bcast32", N.MODEL_BCAST32, lambda: D.bcast32_program(2), 0, 200, 0, 200, 0),
]


@pytest.mark.parametrize("case", FIFO_CASES, ids=[c[0] for c in FIFO_CASES])
def test_src_dst_fifo_strategy_matches_oracle(case, oracle):
    """RandomizationStrategy = SrcDstFIFO (RandomScheduler.scala:702-909): random (src,dst) pair, FIFO within it."""
    _, model, prog, flags, maxm, interval, n, blocked = case
    ext = D.pack_externals(prog())
    eng = D.Engine(D.SchedulerConfig(model, model_flags=flags, blocked_mask=blocked, strategy=1))
    eng.set_externals(ext)
    gpu = eng.fuzz_batch(3, n, maxm, interval, flags=1)
    cpu = oracle.fuzz_batch(model, ext, 3, n, maxm, interval, model_flags=flags, blocked_mask=blocked, flags=1, strategy=1)
    assert_same(gpu, cpu)
    assert (gpu["status"] == 0).all()
    fully = oracle.fuzz_batch(model, ext, 3, min(n, 300), maxm, interval, model_flags=flags, blocked_mask=blocked, flags=1)
    assert (fully["trace_hash"] != cpu["trace_hash"][:len(fully)]).any()       # it really is a different strategy
    gev, gpar, gres = eng.fuzz_trace(11, maxm, interval)
    cev, cpar, cres = oracle.fuzz_trace(model, ext, 11, maxm, interval, model_flags=flags, blocked_mask=blocked, strategy=1)
    assert (gev == cev).all() and (gpar == cpar).all()
    # per-pair FIFO: deliveries of one (src,dst) pair happen in send order
    sends = {}
    for e in cev:
        if e["kind"] == N.EV_MSG_SEND and e["src"] < 32:
            sends.setdefault((int(e["src"]), int(e["dst"])), []).append(int(e["uniq"]))
    for e in cev:
        if e["kind"] == N.EV_MSG_EVENT and e["src"] < 32:
            q = sends[(int(e["src"]), int(e["dst"]))]
            # dropped (partitioned) sends never arrive; everything delivered comes out in order
            while q and q[0] != int(e["uniq"]):
                q.pop(0)
            assert q and q.pop(0) == int(e["uniq"])


def test_user_filter_and_hard_kill(oracle):
    """FullyRandom(userDefinedFilter) — the redraw loop as written (RandomScheduler.scala:666-684) — and
    HardKill -> Scheduler.actorTerminated -> FullyRandom.removeAll (:686-696, :536-547), on the general (warp) engine."""
    rules = [(0b00110, 0b11111, 1 << 5, 0), (0, 0b00001, 1 << 3, N.FRULE_DEADLETTERS)]
    hk = D.raft5_program(client_cmds=3)[:-1] + [D.WaitQuiescence(), D.HardKill(1), D.Send(1, 2, 40), D.Send(2, 2, 41), D.WaitQuiescence(),
                                                D.Start(1), D.Send(1, 1, 0x1F), D.Kill(3), D.WaitQuiescence(), D.HardKill(0), D.WaitQuiescence()]
    for prog, flt, maxm, interval, blocked in [(D.raft5_program(), rules, 50, 5, 0), (hk, [], 90, 7, 0), (hk, rules[:1], 90, 0, 0b00010),
                                               (D.raft5_program(client_cmds=6), rules, 200, 20, 0b01000)]:
        ext = D.pack_externals(prog)
        eng = D.Engine(D.SchedulerConfig(N.MODEL_RAFT5, model_flags=1, blocked_mask=blocked))
        eng.set_user_filter(flt)
        eng.set_externals(ext)
        n = 6000
        gpu = eng.fuzz_batch(1, n, maxm, interval)
        oracle.set_user_filter(flt)
        try:
            cpu = oracle.fuzz_batch(N.MODEL_RAFT5, ext, 1, n, maxm, interval, model_flags=1, blocked_mask=blocked)
            assert_same(gpu, cpu)
            assert (gpu["status"] == 0).all()
            for seed in (3, 77, 1234):
                ev, par, r = eng.fuzz_trace(seed, maxm, interval)
                cev, cpar, _ = oracle.fuzz_trace(N.MODEL_RAFT5, ext, seed, maxm, interval, model_flags=1, blocked_mask=blocked)
                assert (ev == cev).all() and (par == cpar).all()
        finally:
            oracle.set_user_filter([])
    # replay and DPOR refuse HardKill, like the reference's DPOR ("unsuported external event")
    eng = D.Engine(D.SchedulerConfig(N.MODEL_RAFT5, model_flags=1))
    with pytest.raises(D.DemiError):
        eng.dpor_batch([hk[:-1]], 50, 10)
    with pytest.raises(D.DemiError):
        eng.dpor_frontier([e for e in hk if not isinstance(e, D.WaitQuiescence)], eng.frontier_params(50, 10, 4))
