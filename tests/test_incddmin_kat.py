"""Oracle checks for the edit-distance bounded, resumable DPOR configuration and IncrementalDDMin
(BacktrackOrdering.scala:99-173, DPORwHeuristics.scala:128-134 / :1142-1162 / :1193-1242,
minification/IncrementalDeltaDebugging.scala:20-122)."""
import ctypes as C

import numpy as np
import pytest

from demi_b200 import _native as N
from demi_b200 import events as E
from oracle import binding as O


def arvind(oi):
    a = np.asarray(oi, dtype=np.int32)
    f = O.lib().oracle_arvind_distance_of
    f.restype = C.c_uint32
    return f(C.c_void_p(a.ctypes.data), C.c_uint32(len(a)))


def test_arvind_distance_hand_derived():
    # (i) events absent from the original (+1 each) plus (ii) earlier path elements the original orders later
    assert arvind([0, 1, 2, 3]) == 0
    assert arvind([0, 2, -1, 1]) == 2            # one absent, and 2 precedes 1
    assert arvind([3, 2, 1, 0]) == 6             # every pair inverted
    assert arvind([-1, -1]) == 2                 # absent events never count as predecessors
    assert arvind([5, -1, 4, 4]) == 3            # 5>4 twice, one absent; equal indices are not "after"


def scenario(seed_index=1):
    ext_all = E.pack_externals(E.raft5_program())
    dext = ext_all[(ext_all["kind"] == 1) | (ext_all["kind"] == 3)]
    ev, par, r = O.fuzz_trace(N.MODEL_RAFT5, ext_all, 1 + seed_index, 40, 5, model_flags=1)
    assert r["violation"] == 1
    return dext, ev, par, int(r["steps"])


def test_resumable_instance_semantics():
    dext, ev, par, m = scenario(100)             # an execution the seeded DPOR run does not reproduce at once
    seed = O.dpor_seed(ev, par)
    inst = O.DporInstance(N.MODEL_RAFT5, dext, m, 400, seed=seed, arvind=1, prioritize_pending=1, model_flags=1,
                          looking_for=1)
    r0, h0 = inst.test(0)
    # setMaxDistance(0): every head distance is >= 0, so getNext stops after the first interleaving (:1145-1146)
    assert r0["interleavings"] == 1 and r0["violations"] == 0 and r0["heap_left"] > 0
    r1, h1 = inst.test(2)
    # resumed with a non-empty backtrack set: the last trace is analysed again (:1219-1220) — the same races are
    # pushed a second time — and, still capped, an unguided execution follows (:757-759)
    assert r1["interleavings"] >= 1 and r1["heap_left"] > r0["heap_left"]
    inst.close()
    # the uncapped default ordering explores many interleavings from the same seed
    inst = O.DporInstance(N.MODEL_RAFT5, dext, m, 50, seed=seed, model_flags=1, looking_for=1)
    r, h = inst.test(-1)
    assert r["interleavings"] > 1
    assert len(set(h.tolist())) > 1
    inst.close()


def test_found_instance_answers_immediately():
    dext, ev, par, m = scenario(1)
    seed = O.dpor_seed(ev, par)
    inst = O.DporInstance(N.MODEL_RAFT5, dext, m, 100, seed=seed, arvind=1, prioritize_pending=1, model_flags=1,
                          looking_for=1)
    r0, _ = inst.test(0)
    assert r0["violations"] == 1 and r0["interleavings"] == 1      # the seeded run replays the recorded execution
    r1, _ = inst.test(2)
    assert r1["violations"] == 1 and r1["interleavings"] == 0      # "Already have shortestTrace!" (:1197-1201)
    inst.close()


@pytest.mark.parametrize("seed_index,expect", [(1, 6), (10, 6), (77, 8), (154, 6)])
def test_incremental_ddmin_oracle(seed_index, expect):
    dext, ev, par, m = scenario(seed_index)
    seed = O.dpor_seed(ev, par)
    rc, mcs, st = O.incremental_ddmin(N.MODEL_RAFT5, dext, m, 2000, seed, model_flags=1, looking_for=1,
                                      stop_at_size=1, max_max_distance=64)
    assert rc == 0
    size = bin(int(mcs[0])).count("1")
    assert size == expect == st["mcs_sizes"][-1]
    assert st["rounds"] == 6                      # caps 0, 2, 4, 8, 16, 32
    assert list(st["mcs_sizes"]) == sorted(st["mcs_sizes"], reverse=True)
    # verify_mcs (:79-87): a fresh instance on the MCS reproduces the violation under the last cap
    sub = dext[[i for i in range(len(dext)) if (int(mcs[0]) >> i) & 1]]
    inst = O.DporInstance(N.MODEL_RAFT5, sub, m, 2000, seed=seed, arvind=1, prioritize_pending=1, model_flags=1,
                          looking_for=1)
    found = False
    for cap in (0, 2, 4, 8, 16, 32):
        r, _ = inst.test(cap)
        found = found or r["violations"] > 0
    inst.close()
    assert found
