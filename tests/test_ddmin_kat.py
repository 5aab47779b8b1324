"""Known-answer checks for the DDMin / EventDag / STSSched restatement."""
import numpy as np

import demi_b200 as D
from demi_b200 import _native as N


def bits(mask):
    return [i for i in range(64 * len(mask)) if (int(mask[i // 64]) >> (i % 64)) & 1]


def test_split_list_first_chunk_gets_the_extra_element(oracle):
    # split_list([0..6], 2) = [0..3], [4..6]   (minification/Util.scala:16-35)
    assert [oracle.split_first_len(7, 2, 0), oracle.split_first_len(7, 2, 1)] == [4, 3]
    assert [oracle.split_first_len(8, 2, 0), oracle.split_first_len(8, 2, 1)] == [4, 4]
    assert [oracle.split_first_len(1, 2, 0), oracle.split_first_len(1, 2, 1)] == [1, 0]
    assert [oracle.split_first_len(10, 3, k) for k in range(3)] == [4, 3, 3]


def test_ddmin_on_monotone_oracle_returns_exactly_K_and_the_reference_test_order(oracle):
    ext = D.pack_externals([D.Send(0, 1, k) for k in range(8)])
    K = np.array([(1 << 2) | (1 << 5)], dtype=np.uint64)
    rc, mcs, total, iters, log = oracle.ddmin_superset(ext, K)
    assert rc == 0 and bits(mcs) == [2, 5]
    # hand-derived from DeltaDebugging.scala:73-109: [left, right] at each level, first failing split wins,
    # interference recursion left-then-right with the sibling half as remainder
    expect = [
        [0, 1, 2, 3], [4, 5, 6, 7],                  # both halves pass -> interference
        [0, 1, 4, 5, 6, 7], [2, 3, 4, 5, 6, 7],      # ddmin2([0..3], rem=[4..7]): right half fails
        [2, 4, 5, 6, 7],                             # ddmin2([2,3], rem): left fails -> {2}
        [0, 1, 2, 3, 4, 5],                          # ddmin2([4..7], rem=[0..3]): left half [4,5] fails at once
        [0, 1, 2, 3, 4], [0, 1, 2, 3, 5],            # ddmin2([4,5], rem): left passes, right fails -> {5}
    ]
    assert [bits(m) for m in log] == expect
    assert total == len(expect)
    # record_iteration_size(original - pruned) after every test + the fencepost (DeltaDebugging.scala:60, :94)
    assert list(iters) == [8, 8, 8, 8, 6, 5, 3, 3, 2]


def test_ddmin_singleton_and_everything(oracle):
    ext = D.pack_externals([D.Send(0, 1, k) for k in range(5)])
    rc, mcs, total, iters, log = oracle.ddmin_superset(ext, np.array([1 << 3], dtype=np.uint64))
    assert rc == 0 and bits(mcs) == [3]
    rc, mcs, total, iters, log = oracle.ddmin_superset(ext, np.array([0b11111], dtype=np.uint64))
    assert rc == 0 and bits(mcs) == [0, 1, 2, 3, 4]
    # empty K: the first tested half already "violates" all the way down to one atom
    rc, mcs, total, iters, log = oracle.ddmin_superset(ext, np.array([0], dtype=np.uint64))
    assert rc == 0 and len(bits(mcs)) == 1


def test_ddmin_atomic_pairs_are_never_split(oracle):
    # Start(0) .. Kill(0) and Partition(1,2) .. UnPartition(1,2) are single atoms (minification/Util.scala:197-265)
    prog = [D.Start(0), D.Start(1), D.Send(1, 1, 5), D.Kill(0), D.Partition(1, 2), D.Send(1, 1, 6),
            D.UnPartition(1, 2), D.WaitQuiescence(), D.Send(1, 1, 7)]
    ext = D.pack_externals(prog)
    K = np.array([(1 << 3) | (1 << 6)], dtype=np.uint64)       # needs the Kill and the UnPartition
    rc, mcs, total, iters, log = oracle.ddmin_superset(ext, K)
    assert rc == 0
    assert bits(mcs) == [0, 3, 4, 6]                           # ... so their duals come along
    for m in log:
        b = set(bits(m))
        assert (0 in b) == (3 in b) and (4 in b) == (6 in b)
        assert 7 not in b                                       # WaitQuiescence is dropped (RunnerUtils.scala:678-684)
    # Kill without a preceding Start is malformed
    bad = D.pack_externals([D.Kill(0), D.Send(0, 1, 1)])
    rc, *_ = oracle.ddmin_superset(bad, np.array([1], dtype=np.uint64))
    assert rc == -2


def test_projection_as_written(oracle):
    """subsequenceIntersection / filterSends / filterKnownAbsentInternals (EventTrace.scala:290-534)."""
    prog = [D.Start(0), D.Start(1), D.Start(2), D.Send(0, 1, 10), D.Send(1, 1, 11), D.Send(2, 1, 12), D.WaitQuiescence()]
    ext = D.pack_externals(prog)
    ev, par, r = oracle.fuzz_trace(N.MODEL_PINGPONG3, ext, 4, -1, 0)
    full = oracle.full_mask(ext)
    keep = oracle.project(N.MODEL_PINGPONG3, ev, ext, full)
    assert keep.all()                                           # the full subsequence keeps everything
    # drop the 2nd Send: its MsgSend and its delivery disappear; the Pong it caused is an internal
    # event and stays expected (it will be skipped at replay time)
    m = full.copy(); m[0] &= ~np.uint64(1 << 4)
    keep = oracle.project(N.MODEL_PINGPONG3, ev, ext, m)
    dropped = ev[keep == 0]
    assert len(dropped) == 2 and set(dropped["kind"]) == {N.EV_MSG_SEND, N.EV_MSG_EVENT}
    assert (dropped["p0"] == 11).all() and (dropped["type"] == 1).all()
    # drop Start(1): without filterKnownAbsents only the SpawnEvent goes ...
    m = full.copy(); m[0] &= ~np.uint64(1 << 1)
    keep = oracle.project(N.MODEL_PINGPONG3, ev, ext, m)
    assert (keep == 0).sum() == 1 and ev[keep == 0][0]["kind"] == N.EV_SPAWN
    # ... with it, sends from actor 1 and deliveries to actor 1 are pruned too
    keep2 = oracle.project(N.MODEL_PINGPONG3, ev, ext, m, filter_known_absents=True)
    gone = ev[(keep == 1) & (keep2 == 0)]
    assert len(gone) > 0
    for e in gone:
        assert (e["kind"] == N.EV_MSG_SEND and e["src"] == 1) or (e["kind"] == N.EV_MSG_EVENT and (e["dst"] == 1 or e["src"] == 1))


def test_sts_replay_skips_absent_and_reproduces(oracle):
    ext = D.pack_externals(D.raft5_program(client_cmds=3))
    res = oracle.fuzz_batch(N.MODEL_RAFT5, ext, 1, 3000, 50, 5, model_flags=1)
    seed = 1 + int(np.nonzero(res["violation"])[0][0])
    ev, par, r = oracle.fuzz_trace(N.MODEL_RAFT5, ext, seed, 50, 5, model_flags=1)
    full = oracle.full_mask(ext)
    out = oracle.replay_batch(N.MODEL_RAFT5, ev, ext, [full, np.zeros_like(full)], looking_for=int(r["violation"]), model_flags=1)
    n_del = int((ev["kind"] == N.EV_MSG_EVENT).sum())
    assert out[0]["violation"] == r["violation"] and out[0]["delivered"] == n_del and out[0]["ignored"] == 0
    # empty subsequence: deliveries of the pruned external Sends vanish from the expected trace (filterSends),
    # every other expected delivery is skipped because nothing is ever pending (A.7 skip rule)
    n_ext_del = int(((ev["kind"] == N.EV_MSG_EVENT) & ((ev["type"] == 1) | (ev["type"] == 2))).sum())
    assert out[1]["violation"] == 0 and out[1]["delivered"] == 0 and out[1]["ignored"] == n_del - n_ext_del
    # the replayed final state equals the fuzz run's final state: compare through a strict replay too
    strict = oracle.replay_batch(N.MODEL_RAFT5, ev, ext, [full], looking_for=int(r["violation"]), flags=2, model_flags=1)
    assert strict[0]["status"] == 0 and strict[0]["state_hash"] == out[0]["state_hash"]
