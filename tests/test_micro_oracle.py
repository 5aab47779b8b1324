"""Cross-check of the C oracle (oracle/*.c) against the independently written Python restatement of the same Scala
(tests/micro_oracle.py): 10^4 seeds, event by event.  Both are restatements; they share no code, no headers and no
author-time assumptions beyond the Scala sources and DESIGN.md §3's model specification."""
import os

import numpy as np

import demi_b200 as D
from demi_b200 import _native as N
import micro_oracle as M


def to_prog(prog):
    out = []
    for e in prog:
        k = type(e).__name__
        if k in ("Start", "Kill", "HardKill"):
            out.append((k, str(e.a)))
        elif k == "Send":
            out.append(("Send", str(e.a), (e.type, e.p0, e.p1)))
        elif k in ("Partition", "UnPartition"):
            out.append((k, str(e.a), str(e.b)))
        else:
            out.append(("WaitQuiescence",))
    return out


def name_idx(s):
    return 0xFE if s == "Timer" else 0xFF if s == M.DEADLETTERS else int(s)


def flat(events):
    rows = []
    for e in events:
        if e[0] in ("MsgSend", "MsgEvent"):
            _, snd, rcv, (t, p0, p1), uniq, node = e
            rows.append((1 if e[0] == "MsgSend" else 2, name_idx(snd), int(rcv), t, p0, p1, uniq, node))
        elif e[0] == "Spawn":
            rows.append((3, 0xFF, int(e[1]), 0, 0, 0, 0, 0))
        elif e[0] == "Kill":
            rows.append((4, 0xFF, int(e[1]), 0, 0, 0, 0, 0))
        elif e[0] == "HardKill":
            rows.append((9, 0xFF, int(e[1]), 0, 0, 0, 0, 0))
        elif e[0] == "Partition":
            rows.append((5, int(e[1]), int(e[2]), 0, 0, 0, 0, 0))
        elif e[0] == "UnPartition":
            rows.append((6, int(e[1]), int(e[2]), 0, 0, 0, 0, 0))
        elif e[0] == "BeginWaitQuiescence":
            rows.append((7, 0xFF, 0xFF, 0, 0, 0, 0, 0))
        else:
            rows.append((8, 0xFF, 0xFF, 0, 0, 0, 0, 0))
    return rows


def run_micro(model, prog, seed, maxm, interval, flags, rules=(), strategy=0):
    if model == N.MODEL_RAFT5:
        fresh = lambda name: M.RaftActor(int(name), flags)
        actors = {str(i): fresh(str(i)) for i in range(5)}
        inv = M.raft_invariant
    else:
        fresh = lambda name: M.PingPongActor(int(name))
        actors = {str(i): fresh(str(i)) for i in range(3)}
        inv = M.pingpong_invariant(flags)
    ext_types = (1, 2) if model == N.MODEL_RAFT5 else (1,)

    def user_filter(snd, rcv, msg):                                  # the closure the rules stand for
        for src_mask, dst_mask, type_mask, fl in rules:
            src_ok = ((src_mask >> int(snd)) & 1) if snd.isdigit() else (fl & 1)
            if src_ok and (dst_mask >> int(rcv)) & 1 and (type_mask >> msg[0]) & 1:
                return False
        return True
    cls = M.SrcDstFifoExecution if strategy == 1 else M.Execution
    ex = cls(actors, to_prog(prog), seed, maxm, interval, inv, lambda m: m[0] in ext_types,
             user_filter=user_filter if rules else None, fresh_actor=fresh)
    v = ex.run()
    return ex, (v or 0)


def compare(oracle, model, prog, seeds, maxm, interval, flags, full_trace_every=1, rules=()):
    ext = D.pack_externals(prog)
    oracle.set_user_filter(rules)
    try:
        return _compare(oracle, model, prog, ext, seeds, maxm, interval, flags, full_trace_every, rules)
    finally:
        oracle.set_user_filter([])


def _compare(oracle, model, prog, ext, seeds, maxm, interval, flags, full_trace_every, rules):
    res = oracle.fuzz_batch(model, ext, seeds[0], len(seeds), maxm, interval, model_flags=flags)
    n_viol = 0
    for k, seed in enumerate(seeds):
        ex, v = run_micro(model, prog, seed, maxm, interval, flags, rules)
        r = res[k]
        assert (int(r["violation"]), int(r["steps"]), int(r["n_nodes"]), int(r["n_events"]), int(r["max_pending"]), int(r["status"])) == \
               (v, ex.messagesScheduledSoFar, ex.depTracker.next_id, len(ex.events), ex.max_pending, 0), seed
        n_viol += bool(v)
        if k % full_trace_every == 0:
            ev, par, _ = oracle.fuzz_trace(model, ext, seed, maxm, interval, model_flags=flags)
            rows = flat(ex.events)
            assert len(rows) == len(ev), seed
            got = [(int(e["kind"]), int(e["src"]), int(e["dst"]), int(e["type"]), int(e["p0"]), int(e["p1"]), int(e["uniq"]), int(e["node"]))
                   for e in ev]
            assert got == rows, (seed, next(i for i in range(len(rows)) if got[i] != rows[i]))
            assert [int(x) for x in par] == [ex.depTracker.parent_of[i] for i in range(ex.depTracker.next_id)], seed
    return n_viol


def test_raft5_bench_workload_10k_seeds(oracle):
    """The headline workload (raft5, depth 50, invariant every 5, double-vote bug): verdicts + counters on 10^4 seeds,
    the whole EventTrace and DepTracker tree on every 10th."""
    n = int(os.environ.get("DEMI_MICRO_SEEDS", "10000"))
    n_viol = compare(oracle, N.MODEL_RAFT5, D.raft5_program(), list(range(1, n + 1)), 50, 5, 1, full_trace_every=10)
    assert n_viol > n // 50                                          # the seeded bug is found (~5 % of the prefixes)


def test_raft5_with_kills_partitions_and_client_commands(oracle):
    prog = D.raft5_program(client_cmds=4)[:-1] + [D.WaitQuiescence(), D.Partition(0, 1), D.Kill(2), D.Send(3, 2, 9), D.WaitQuiescence(),
                                                  D.UnPartition(0, 1), D.Start(2), D.Send(0, 2, 11), D.WaitQuiescence()]
    compare(oracle, N.MODEL_RAFT5, prog, list(range(1, 401)), 120, 7, 3, full_trace_every=4)
    compare(oracle, N.MODEL_RAFT5, prog, list(range(1000, 1200)), -1 if False else 300, 0, 2, full_trace_every=4)


def test_user_filter_and_hard_kill(oracle):
    """FullyRandom(userDefinedFilter) as written (redraw while more than one element is left, rejected draws go back
    after the loop) and HardKill -> actorTerminated -> removeAll + a fresh instance on the next Start."""
    rules = [(0b00110, 0b11111, 1 << 5, 0),                          # VoteReplies from nodes 1, 2 are held back ...
             (0, 0b00001, 1 << 3, 1)]                                # ... and node 0's election tick (a timer: deadLetters)
    compare(oracle, N.MODEL_RAFT5, D.raft5_program(), list(range(1, 801)), 50, 5, 1, full_trace_every=4, rules=rules)
    prog = D.raft5_program(client_cmds=3)[:-1] + [D.WaitQuiescence(), D.HardKill(1), D.Send(1, 2, 40), D.Send(2, 2, 41), D.WaitQuiescence(),
                                                  D.Start(1), D.Send(1, 1, 0x1F), D.Kill(3), D.WaitQuiescence(), D.HardKill(0), D.WaitQuiescence()]
    compare(oracle, N.MODEL_RAFT5, prog, list(range(1, 601)), 90, 7, 1, full_trace_every=3)
    compare(oracle, N.MODEL_RAFT5, prog, list(range(1, 301)), 90, 0, 3, full_trace_every=3, rules=rules[:1])


def test_pingpong3_config0(oracle):
    """BASELINE.json configs[0]: 3-actor ping-pong, 100 external messages, seed = 1 (and its neighbours)."""
    compare(oracle, N.MODEL_PINGPONG3, D.pingpong3_program(100), list(range(1, 301)), -1, 0, 0, full_trace_every=3)
    compare(oracle, N.MODEL_PINGPONG3, D.pingpong3_program(20), list(range(1, 2001)), -1, 3, 1 | (4 << 8), full_trace_every=20)


def test_ddmin2_test_sequence_matches_the_c_oracle(oracle):
    """DDMin.ddmin2 over a monotone oracle ("violates iff the subsequence contains K"): same tests, in the same order."""
    rng = np.random.default_rng(0)
    for trial in range(40):
        n = int(rng.integers(4, 28))
        prog = [D.Start(a) for a in range(3)] + [D.Send(k % 3, 1, k) for k in range(n)]
        if trial % 3 == 0:
            prog += [D.Partition(0, 1), D.Send(2, 1, 99), D.UnPartition(0, 1), D.Kill(2)]
        ext = D.pack_externals(prog)
        K = set(int(x) for x in rng.choice(np.arange(3, 3 + n), size=int(rng.integers(1, 4)), replace=False))
        Kmask = np.zeros(1, dtype=np.uint64)
        for i in K:
            Kmask[0] |= np.uint64(1 << i)
        rc, mcs, total, iters, log = oracle.ddmin_superset(ext, Kmask)
        assert rc == 0
        events = to_prog(prog)
        index = {id(e): i for i, e in enumerate(events)}
        dd = M.DDMin(lambda sub: K <= set(index[id(e)] for e in sub))
        got = dd.minimize(events)
        assert sorted(i for i, _ in got) == [i for i in range(len(prog)) if (int(mcs[0]) >> i) & 1]
        assert dd.total_replays == total
        assert [sum(1 << i for i in t) for t in dd.tests] == [int(m[0]) for m in log]


def _replay_case(oracle, model, prog, seed, maxm, interval, flags, n_masks, rng, fka_modes=(0, 1), with_wait_quiescence=False):
    """One recorded execution, many external subsequences: STSScheduler.test restated in Python (straight from
    STSScheduler.scala / EventTrace.scala) against the C oracle — verdict, delivered / ignored counts and the whole
    recorded EventTrace of every replay."""
    ext = D.pack_externals(prog)
    events = to_prog(prog)
    ex, v = run_micro(model, prog, seed, maxm, interval, flags)
    ev, par, r = oracle.fuzz_trace(model, ext, seed, maxm, interval, model_flags=flags)
    assert [tuple(int(e[f]) for f in ("kind", "src", "dst", "type", "p0", "p1", "uniq")) for e in ev] == [row[:7] for row in flat(ex.events)]
    code = int(r["violation"])
    n_ext = len(ext)
    full = np.asarray(oracle.full_mask(ext, drop_wait_quiescence=not with_wait_quiescence), dtype=np.uint64).reshape(-1)
    mw = oracle.mask_words(n_ext)
    ext_types = (1, 2) if model == N.MODEL_RAFT5 else (1,)
    checked = 0
    for t in range(n_masks):
        keep = rng.random(n_ext) < rng.choice([0.95, 0.8, 0.5, 0.25])
        if t == 0:
            keep[:] = True
        mask = np.zeros(mw, dtype=np.uint64)
        for i in range(n_ext):
            if keep[i]:
                mask[i >> 6] |= np.uint64(1) << np.uint64(i & 63)
        mask &= full
        subseq = [i for i in range(n_ext) if (int(mask[i >> 6]) >> (i & 63)) & 1]
        for fka in fka_modes:
            if model == N.MODEL_RAFT5:
                actors = {str(i): M.RaftActor(i, flags) for i in range(5)}
                inv = M.raft_invariant
            else:
                actors = {str(i): M.PingPongActor(i) for i in range(3)}
                inv = M.pingpong_invariant(flags)
            sts = M.STSReplay(actors, ex.events, events, inv, lambda m: m[0] in ext_types, filter_known_absents=bool(fka),
                              looking_for=code or None)
            got = sts.test(subseq) or 0
            res, rec = oracle.replay_trace(model, ev, ext, mask, looking_for=code, flags=fka, model_flags=flags)
            assert int(res["status"]) == 0
            assert (got, sts.delivered, sts.ignored) == (int(res["violation"]), int(res["delivered"]), int(res["ignored"])), (seed, t, fka)
            rows = [row[:7] for row in flat(sts.events)]
            crow = [tuple(int(e[f]) for f in ("kind", "src", "dst", "type", "p0", "p1", "uniq")) for e in rec]
            assert rows == crow, (seed, t, fka, next((i for i in range(min(len(rows), len(crow))) if rows[i] != crow[i]), -1))
            checked += 1
    return code, checked


def test_sts_replay_of_external_subsequences_matches_the_c_oracle(oracle):
    rng = np.random.default_rng(5)
    n_checked = n_viol = 0
    res = oracle.fuzz_batch(N.MODEL_RAFT5, D.pack_externals(D.raft5_program(client_cmds=3)), 1, 600, 60, 5, model_flags=1)
    seeds = [1 + int(i) for i in np.nonzero(res["violation"])[0][:6]] + [2, 3]
    for seed in seeds:
        code, c = _replay_case(oracle, N.MODEL_RAFT5, D.raft5_program(client_cmds=3), seed, 60, 5, 1, 40, rng)
        n_checked += c; n_viol += bool(code)
    prog = D.raft5_program(client_cmds=4)[:-1] + [D.WaitQuiescence(), D.Partition(0, 1), D.Kill(2), D.Send(3, 2, 9), D.WaitQuiescence(),
                                                  D.UnPartition(0, 1), D.Start(2), D.Send(0, 2, 11), D.WaitQuiescence()]
    for seed in (1, 2, 3, 4):
        _, c = _replay_case(oracle, N.MODEL_RAFT5, prog, seed, 120, 7, 3, 40, rng)
        n_checked += c
    _, c = _replay_case(oracle, N.MODEL_RAFT5, prog, 5, 120, 7, 3, 30, rng, with_wait_quiescence=True)   # WaitQuiescence left in the subsequence
    n_checked += c
    for seed in (1, 2, 3):
        _, c = _replay_case(oracle, N.MODEL_PINGPONG3, D.pingpong3_program(20), seed, -1, 3, 1 | (4 << 8), 30, rng)
        n_checked += c
    assert n_viol >= 4 and n_checked >= 1200


def _dpor_case(oracle, model, prog, flags, max_messages, budget, stop_if_found, looking_for=0):
    ext = D.pack_externals(prog)
    rc, r, viol, hashes = oracle.dpor_search(model, ext, max_messages, budget, looking_for=looking_for,
                                             stop_if_found=1 if stop_if_found else 0, model_flags=flags)
    assert rc == 0 and int(r["status"]) == 0
    if model == N.MODEL_RAFT5:
        make = lambda: {str(i): M.RaftActor(i, flags) for i in range(5)}
        inv = M.raft_invariant
    else:
        make = lambda: {str(i): M.PingPongActor(i) for i in range(3)}
        inv = M.pingpong_invariant(flags)
    s = M.DPORSearch(make, to_prog(prog), inv, max_messages, stop_if_found=stop_if_found, looking_for=looking_for or None)
    exhausted = s.search(budget)

    def sched_hash(trace):
        h = 0
        for i, u in enumerate(trace):
            if i == 0:
                continue
            snd, rcv, (t, p0, p1) = s.event[u]
            h = (h + M.hash6(name_idx(snd) | (int(rcv) << 8) | (t << 16), p0, p1, i, 0, 0)) & ((1 << 64) - 1)
        return h
    assert len(s.traces) == int(r["interleavings"]), (len(s.traces), int(r["interleavings"]))
    assert [sched_hash(t) for t in s.traces] == [int(h) for h in hashes]                      # the same schedules, in the same order
    assert sum(len(t) - 1 for t in s.traces) == int(r["deliveries"])
    assert s.races == int(r["races"]) and s.next_id == int(r["n_nodes"]) and len(s.explored) == int(r["n_explored"])
    assert [(k, v) for k, v in s.violations] == [(int(x["interleaving"]), int(x["code"])) for x in viol]
    assert bool(exhausted) == bool(r["exhausted"])
    return len(s.traces), len(s.violations)


def test_dpor_search_matches_the_c_oracle(oracle):
    """DPORwHeuristics restated in Python straight from the Scala (graph of Uniques with child reuse, nextTrace matching,
    the race scan with getCommonPrefix, ExploredTacker, getNext) against the C oracle: the same schedules in the same
    order, the same counters, on exhaustive and budgeted raft5 / pingpong3 searches, with and without a seeded bug."""
    rng = np.random.default_rng(11)
    total = viols = 0
    for trial in range(14):
        starts = [int(a) for a in rng.permutation(5)]
        boots = [int(a) for a in rng.permutation(5)[:int(rng.integers(2, 6))]]
        prog = [D.Start(a) for a in starts] + [D.Send(a, 1, 0x1F) for a in boots]
        if trial % 3 == 0:
            prog.append(D.Send(int(rng.integers(0, 5)), 2, int(rng.integers(1, 50))))
        flags = [0, 1, 3][trial % 3]
        maxm = int(rng.integers(14, 30))
        n, v = _dpor_case(oracle, N.MODEL_RAFT5, prog, flags, maxm, 400, stop_if_found=(trial % 2 == 0))
        total += n; viols += v
    n, v = _dpor_case(oracle, N.MODEL_RAFT5, [D.Start(a) for a in range(5)] + [D.Send(a, 1, 0x1F) for a in range(5)], 1, 60, 150, False)
    total += n; viols += v
    pprog = [D.Start(0), D.Start(1), D.Start(2), D.Send(2, 1, 0), D.Send(2, 1, 1), D.Send(0, 1, 2)]
    n, v = _dpor_case(oracle, N.MODEL_PINGPONG3, pprog, 1 | (2 << 8), 40, 500, False, looking_for=7)
    total += n; viols += v
    assert total > 300 and viols > 0


KIND_NAME = {1: "MsgSend", 2: "MsgEvent", 3: "Spawn", 4: "Kill", 5: "Partition", 6: "UnPartition", 7: "BeginWaitQuiescence", 8: "Quiescence"}


def unflat(ev):
    """C event records -> the tuples the micro-oracle works on."""
    nm = lambda x: "Timer" if x == 0xFE else M.DEADLETTERS if x == 0xFF else str(x)
    out = []
    for e in ev:
        k = KIND_NAME[int(e["kind"])]
        if k in ("MsgSend", "MsgEvent"):
            out.append((k, nm(int(e["src"])), str(int(e["dst"])), (int(e["type"]), int(e["p0"]), int(e["p1"])), int(e["uniq"]), 0))
        elif k in ("Spawn", "Kill"):
            out.append((k, str(int(e["dst"]))))
        elif k in ("Partition", "UnPartition"):
            out.append((k, str(int(e["src"])), str(int(e["dst"]))))
        else:
            out.append((k,))
    return out


def test_internal_minimization_matches_the_c_oracle(oracle):
    """STSSchedMinimizer + LeftToRightOneAtATime restated in Python (on top of the Python STSScheduler) against
    oracle_internal_minimize: the minimized schedule, the number of replays, the size series, the unignorable count."""
    cases = 0
    prog = D.raft5_program(client_cmds=2)
    ext = D.pack_externals(prog)
    res = oracle.fuzz_batch(N.MODEL_RAFT5, ext, 1, 400, 40, 5, model_flags=1)
    for i in np.nonzero(res["violation"])[0][:4]:
        seed = 1 + int(i)
        ev, par, r = oracle.fuzz_trace(N.MODEL_RAFT5, ext, seed, 40, 5, model_flags=1)
        code = int(r["violation"])
        full = np.asarray(oracle.full_mask(ext), dtype=np.uint64).reshape(-1)
        rr, vtrace = oracle.replay_trace(N.MODEL_RAFT5, ev, ext, full, looking_for=code, model_flags=1)
        if int(rr["violation"]) != code:
            continue
        mcs_idx = [k for k in range(len(ext)) if (int(full[k >> 6]) >> (k & 63)) & 1]
        mcs_prog = [prog[k] for k in mcs_idx]
        mcs_ext = D.pack_externals(mcs_prog)
        rc, ctrace, total, sizes, unig = oracle.internal_minimize(N.MODEL_RAFT5, vtrace, mcs_ext, code, model_flags=1)
        assert rc == 0
        mcs_events = to_prog(mcs_prog)
        is_ext = lambda m: m[0] in (1, 2)

        def test(candidate):
            actors = {str(k): M.RaftActor(k, 1) for k in range(5)}
            sts = M.STSReplay(actors, candidate, mcs_events, M.raft_invariant, is_ext, looking_for=code)
            return sts.events if sts.test(list(range(len(mcs_events)))) == code else None
        verified = unflat(vtrace)
        sm = M.STSSchedMinimizer(mcs_events, verified, code, M.LeftToRightOneAtATime(verified, is_ext), test)
        final = sm.minimize()
        assert (sm.total_replays, sm.internal_sizes, sm.strategy.unignorable) == (total, [int(x) for x in sizes], unig), seed
        assert [row[:7] for row in flat(final)] == [tuple(int(e[f]) for f in ("kind", "src", "dst", "type", "p0", "p1", "uniq")) for e in ctrace], seed
        # SrcDstFIFORemoval (OneAtATimeRemoval.scala:139-251)
        rc, ftrace, ftotal, fsizes, funig = oracle.internal_minimize(N.MODEL_RAFT5, vtrace, mcs_ext, code, model_flags=1, flags=N.IM_SRC_DST_FIFO)
        assert rc == 0
        sf = M.STSSchedMinimizer(mcs_events, verified, code, M.SrcDstFIFORemoval(verified, is_ext), test)
        ffinal = sf.minimize()
        assert (sf.total_replays, sf.internal_sizes, sf.strategy.unignorable) == (ftotal, [int(x) for x in fsizes], funig), seed
        assert [row[:7] for row in flat(ffinal)] == [tuple(int(e[f]) for f in ("kind", "src", "dst", "type", "p0", "p1", "uniq")) for e in ftrace], seed
        cases += 1
    assert cases >= 2


def test_resumable_edit_distance_dpor_matches_the_c_oracle(oracle):
    """The configuration IncrementalDDMin drives (seeded dependency graph + initial trace, prioritizePendingUponDivergence,
    ArvindDistanceOrdering, setMaxDistance, repeated test() calls on one instance), restated in Python, against
    oracle_dpor_open / oracle_dpor_test: per call the same schedules, the same verdict, the same graph and history sizes."""
    nm = lambda x: M.DEADLETTERS if x == 0xFF else str(x)
    ext_all = D.pack_externals(D.raft5_program())
    dprog = [e for e in D.raft5_program() if type(e).__name__ in ("Start", "Send")]
    dext = D.pack_externals(dprog)
    checked = 0
    for seed_index, caps, arv in ((100, [0, 2, 4, 8, -1], 1), (1, [0, 2, -1], 1), (7, [0, 0, 2, 4, 16], 1), (100, [-1], 0), (33, [0, 4, -1, -1], 1)):
        ev, par, r = oracle.fuzz_trace(N.MODEL_RAFT5, ext_all, 1 + seed_index, 40, 5, model_flags=1)
        m = int(r["steps"])
        nodes, trace = oracle.dpor_seed(ev, par)
        inst = oracle.DporInstance(N.MODEL_RAFT5, dext, m, 300, seed=(nodes, trace), arvind=arv, prioritize_pending=1, model_flags=1,
                                   looking_for=1)
        init_nodes = [(nm(int(w[0]) & 0xFF), str((int(w[0]) >> 8) & 0xFF), ((int(w[0]) >> 16) & 0xFF, int(w[1]), int(w[2])), int(w[3])) for w in nodes]
        py = M.ResumableDPORInstance(lambda: {str(i): M.RaftActor(i, 1) for i in range(5)}, to_prog(dprog), M.raft_invariant, m,
                                     init_nodes=init_nodes, init_trace=[int(x) for x in trace], arvind=bool(arv), prioritize_pending=True,
                                     stop_if_found=True, looking_for=1)

        def sched_hash(tr):
            h = 0
            for i, u in enumerate(tr):
                if i:
                    snd, rcv, (t, p0, p1) = py.event[u]
                    h = (h + M.hash6(name_idx(snd) | (int(rcv) << 8) | (t << 16), p0, p1, i, 0, 0)) & ((1 << 64) - 1)
            return h
        for cap in caps:
            races_before = py.races
            rc, hashes = inst.test(cap)
            ran, found = py.test(cap, 300)
            assert int(rc["status"]) == 0
            assert (len(ran), bool(found)) == (int(rc["interleavings"]), int(rc["violations"]) > 0), (seed_index, cap)
            assert [sched_hash(t) for t in ran] == [int(h) for h in hashes], (seed_index, cap)
            assert (py.races - races_before, py.next_id, len(py.explored)) == (int(rc["races"]), int(rc["n_nodes"]), int(rc["n_explored"])), (seed_index, cap)
            checked += 1
        inst.close()
    assert checked >= 15


def test_incremental_ddmin_over_resumable_dpor_matches_the_c_oracle(oracle):
    """IncrementalDDMin (IncrementalDeltaDebugging.scala:20-88) over ResumableDPOR (:90-122): the Python DDMin drives one
    Python DPOR instance per external subsequence with growing distance caps; MCS, total replays, rounds and the number
    of instances equal oracle_incremental_ddmin's."""
    nm = lambda x: M.DEADLETTERS if x == 0xFF else str(x)
    ext_all = D.pack_externals(D.raft5_program())
    dprog = [e for e in D.raft5_program() if type(e).__name__ in ("Start", "Send")]
    dext = D.pack_externals(dprog)
    devents = to_prog(dprog)
    done = 0
    for seed_index in (1, 7):
        ev, par, r = oracle.fuzz_trace(N.MODEL_RAFT5, ext_all, 1 + seed_index, 40, 5, model_flags=1)
        if int(r["violation"]) != 1:
            continue
        m = int(r["steps"])
        nodes, trace = oracle.dpor_seed(ev, par)
        rc, cmcs, info = oracle.incremental_ddmin(N.MODEL_RAFT5, dext, m, 60, (nodes, trace), max_max_distance=8, stop_at_size=1,
                                                  looking_for=1, model_flags=1)
        assert rc == 0
        init_nodes = [(nm(int(w[0]) & 0xFF), str((int(w[0]) >> 8) & 0xFF), ((int(w[0]) >> 16) & 0xFF, int(w[1]), int(w[2])), int(w[3])) for w in nodes]
        instances = {}
        state = {"dist": 0}

        def test(sub):                                               # ResumableDPOR.test
            key = tuple(id(e) for e in sub)
            if key not in instances:
                instances[key] = M.ResumableDPORInstance(lambda: {str(i): M.RaftActor(i, 1) for i in range(5)}, list(sub), M.raft_invariant, m,
                                                         init_nodes=init_nodes, init_trace=[int(x) for x in trace], arvind=True,
                                                         prioritize_pending=True, stop_if_found=True, looking_for=1)
            return instances[key].test(state["dist"], 60)[1]
        current = list(enumerate(devents))
        total = rounds = 0
        sizes = []
        while state["dist"] < 8 and len(current) > 1:
            dd = M.DDMin(test)
            current = dd.ddmin2(current, [])
            total += dd.total_replays
            rounds += 1
            sizes.append(len(current))
            state["dist"] = 2 if state["dist"] == 0 else state["dist"] << 1
        mask = sum(1 << i for i, _ in current)
        assert (mask, total, rounds, len(instances)) == (int(cmcs[0]), info["total_replays"], info["rounds"], info["instances"]), seed_index
        assert sizes == [int(x) for x in info["mcs_sizes"]]
        done += 1
    assert done >= 1


def test_provenance_pruning_matches_the_c_oracle(oracle):
    """ProvenanceTracker.pruneConcurrentEvents restated in Python (closure by plain reachability) against the C oracle's
    literal restatement (topological sweep): the same deliveries are kept."""
    ext = D.pack_externals(D.raft5_program())
    res = oracle.fuzz_batch(N.MODEL_RAFT5, ext, 1, 2000, 50, 5, model_flags=1)
    viol = np.nonzero(res["violation"])[0][:25]
    assert len(viol) >= 10
    kept_any = 0
    for i in viol:
        seed = 1 + int(i)
        ev, par, r = oracle.fuzz_trace(N.MODEL_RAFT5, ext, seed, 50, 5, model_flags=1)
        keep, out = oracle.fuzz_provenance(N.MODEL_RAFT5, ext, seed, 50, 5, 1, model_flags=1)
        assert int(out["status"]) == 0
        affected = [str(a) for a in range(5) if (int(out["affected_mask"]) >> a) & 1]
        trace = [(0, "null")] + [(int(e["node"]), str(int(e["dst"]))) for e in ev if int(e["kind"]) == 2]
        parent_of = {k: (int(par[k]) if k else None) for k in range(len(par))}
        pt = M.ProvenanceTracker(trace, parent_of)
        kept = pt.pruneConcurrentEvents(affected)
        assert kept == [t for t in range(len(trace)) if (int(keep[t >> 6]) >> (t & 63)) & 1], seed
        assert len(kept) == int(out["n_kept"]) and len(trace) == int(out["n_trace"])
        kept_any += bool(kept)
    assert kept_any >= 10


def test_src_dst_fifo_strategy_matches_the_c_oracle(oracle):
    """RandomScheduler over SrcDstFIFO (RandomScheduler.scala:702-870): verdicts, counters and the whole EventTrace."""
    for prog, maxm, interval, flags, n in ((D.raft5_program(), 50, 5, 1, 300), (D.raft5_program(client_cmds=3), 120, 7, 3, 150)):
        ext = D.pack_externals(prog)
        res = oracle.fuzz_batch(N.MODEL_RAFT5, ext, 1, n, maxm, interval, model_flags=flags, strategy=1)
        for k in range(n):
            seed = 1 + k
            ex, v = run_micro(N.MODEL_RAFT5, prog, seed, maxm, interval, flags, strategy=1)
            r = res[k]
            assert (int(r["violation"]), int(r["steps"]), int(r["n_nodes"]), int(r["n_events"]), int(r["max_pending"]), int(r["status"])) == \
                   (v, ex.messagesScheduledSoFar, ex.depTracker.next_id, len(ex.events), ex.max_pending, 0), seed
            if k % 10 == 0:
                ev, par, _ = oracle.fuzz_trace(N.MODEL_RAFT5, ext, seed, maxm, interval, model_flags=flags, strategy=1)
                got = [(int(e["kind"]), int(e["src"]), int(e["dst"]), int(e["type"]), int(e["p0"]), int(e["p1"]), int(e["uniq"]), int(e["node"])) for e in ev]
                assert got == flat(ex.events), seed
