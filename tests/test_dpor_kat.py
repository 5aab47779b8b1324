"""Hand-derived known-answer checks for the oracle's DPORwHeuristics restatement
(schedulers/DPORwHeuristics.scala:421-648, :1020-1185).  Model: pingpong3 — Ping(k) to X makes X send Pong(k) to
(X+1)%3.  Actor 2 is never started, so messages to it are discarded when scheduled (:626-635).  Canonical
divergent choice: the non-empty (snd,rcv) queue with the lowest index, actor senders before deadLetters."""
from demi_b200 import _native as N
from demi_b200 import events as E
from oracle import binding as O


def run(prog):
    rc, r, viol, h = O.dpor_search(N.MODEL_PINGPONG3, E.pack_externals(prog), 20, 50)
    assert rc == 0 and r["status"] == 0
    return r, h


def test_two_receivers_one_race():
    """a = Ping->A, b = Ping->B; c = Pong A->B (child of a), d = Pong B->C (child of b, discarded).
    Execution 1: a, c, b.  Only b and c share a receiver; neither descends from the other, their common prefix is
    the root (branch 0).  getNext pops (b, c): nextTrace = [root, a, b] -> execution 2: a, b, c.  Its analysis finds
    the same pair again; (c, b) is already explored, so nothing is enqueued: "Tutto finito"."""
    r, h = run([E.Start(0), E.Start(1), E.Send(0, 1, 0), E.Send(1, 1, 1)])
    assert r["interleavings"] == 2 and len(set(h.tolist())) == 2
    assert r["deliveries"] == 6                      # 3 per execution: the message to actor 2 is never delivered
    assert r["races"] == 2 and r["n_explored"] == 2  # {(c, b), (b, c)}
    assert r["n_nodes"] == 5                         # root, a, b, c, d — ids are reused across executions (:773-801)
    assert r["heap_left"] == 0 and r["exhausted"] == 1 and r["violations"] == 0


def test_one_receiver_two_races():
    """a, b = Ping->A; c = Pong A->B (child of a), d = Pong A->B (child of b).
    Execution 1: a, c, b, d — races (a, b) and (c, d), both branch 0.  Oldest first: (b, a) is taken,
    nextTrace = [root, c, b]; c is not pending yet, so the execution diverges to a, then follows b, then c, d:
    execution 2 = a, b, c, d.  It re-finds (a, b) (filtered, (b, a) explored) and finds (c, d) again, enqueued a
    second time.  The older (d, c) key is taken: nextTrace = [root, a, b, d] -> execution 3 = a, b, d, c.  Its two
    races are both filtered, and the remaining key's pair is explored: done."""
    r, h = run([E.Start(0), E.Start(1), E.Send(0, 1, 0), E.Send(0, 1, 1)])
    assert r["interleavings"] == 3 and len(set(h.tolist())) == 3
    assert r["deliveries"] == 12
    assert r["races"] == 6                           # 2 per execution
    assert r["n_explored"] == 4                      # (a,b), (c,d), (b,a), (d,c)
    assert r["n_nodes"] == 5 and r["heap_left"] == 0 and r["exhausted"] == 1


def test_budget_and_depth_bound():
    prog = [E.Start(0), E.Start(1), E.Start(2)] + [E.Send(k % 3, 1, k) for k in range(6)]
    rc, r, _, h = O.dpor_search(N.MODEL_PINGPONG3, E.pack_externals(prog), 30, 7)
    assert rc == 0 and r["interleavings"] == 7 and r["budget_exhausted"] == 1 and r["exhausted"] == 0
    # the depth gate is `currentDepth < stop_at_depth` with currentDepth = depth(parent) + 1 (:278-282, :832):
    # bound 1 enqueues nothing at all, bound 2 only the externals (children of the root)
    rc, r1, _, _ = O.dpor_search(N.MODEL_PINGPONG3, E.pack_externals(prog), 30, 50, depth_bound=1)
    assert rc == 0 and r1["interleavings"] == 1 and r1["deliveries"] == 0 and r1["n_nodes"] == 7
    rc, r2, _, _ = O.dpor_search(N.MODEL_PINGPONG3, E.pack_externals(prog), 30, 50, depth_bound=2)
    assert rc == 0 and r2["deliveries"] == 6 * r2["interleavings"] and r2["exhausted"] == 1
