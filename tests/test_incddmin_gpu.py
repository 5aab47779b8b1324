"""GPU parity for the edit-distance bounded, resumable DPOR configuration (demi_dpor_batch_ex) and
IncrementalDDMin (demi_incremental_ddmin) against the oracle's stateful restatement."""
import numpy as np
import pytest

import demi_b200 as D
from demi_b200 import _native as N
from oracle import binding as O

pytestmark = pytest.mark.gpu

CAPS = [0, 2, 4, 8, 16, -1]


@pytest.fixture(scope="module")
def world():
    eng = D.Engine(D.SchedulerConfig(N.MODEL_RAFT5, model_flags=1))
    prog = D.raft5_program()
    eng.set_externals(prog)
    ext_all = D.pack_externals(prog)
    dext = ext_all[(ext_all["kind"] == 1) | (ext_all["kind"] == 3)]
    res = eng.fuzz_batch(1, 3000, 40, 5)
    viol = np.nonzero(res["violation"] == 1)[0]
    recs = {}
    for i in viol[:12]:
        ev, par, r = eng.fuzz_trace(1 + int(i), 40, 5)
        recs[int(i)] = (ev, par, int(r["steps"]))
    return eng, dext, recs


def oracle_history(dext, ev, par, m, caps, arvind, prio, budget, seeded=True):
    inst = O.DporInstance(N.MODEL_RAFT5, dext, m, budget, seed=O.dpor_seed(ev, par) if seeded else None, arvind=arvind,
                          prioritize_pending=prio, model_flags=1, looking_for=1)
    tot = {"interleavings": 0, "deliveries": 0, "races": 0}
    hashes, last = [], None
    for c in caps:
        r, h = inst.test(c)
        for k in tot:
            tot[k] += int(r[k])
        hashes += h.tolist()
        last = r
    inst.close()
    return tot, hashes, last


@pytest.mark.parametrize("arvind,prio,seeded", [(1, 1, True), (0, 1, True), (0, 0, True), (0, 0, False), (0, 1, False)])
def test_instances_match_oracle(world, arvind, prio, seeded):
    eng, dext, recs = world
    rng = np.random.default_rng(5)
    programs, caps, meta = [], [], []
    budget = 60
    # every search of a launch shares the seed execution: vary the subsequence and the cap history
    ev, par, steps = list(recs.values())[0]
    for t in range(24):
        keep = np.ones(len(dext), dtype=bool)
        if t:
            keep[rng.integers(0, len(dext), size=int(rng.integers(0, 4)))] = False
        hist = CAPS[:int(rng.integers(1, len(CAPS) + 1))]
        if not arvind:
            hist = [c if c < 0 else 0 for c in hist]       # without the distance ordering every key has distance 0
        programs.append(dext[keep]); caps.append(hist)
    flags = (N.DF_ARVIND_ORDERING if arvind else 0) | (N.DF_PRIORITIZE_PENDING if prio else 0)
    res, hashes = eng.dpor_batch_ex(programs, steps, budget, seed=(ev, par) if seeded else None, flags=flags, caps=caps,
                                    looking_for=1, want_hashes=True, heap_cap=1 << 16)
    found = 0
    for t in range(len(programs)):
        tot, oh, last = oracle_history(programs[t], ev, par, steps, caps[t], arvind, prio, budget, seeded)
        r = res[t]
        assert r["status"] == 0 == last["status"]
        for k in tot:
            assert int(r[k]) == tot[k], (t, k)
        for k in ("n_nodes", "n_explored", "heap_left"):
            assert int(r[k]) == int(last[k]), (t, k)
        assert hashes[t][:len(oh)].tolist() == oh, t
        found += int(r["violations"] > 0)
    if seeded:
        assert found > 0          # the seeded instance on the full sequence replays the recorded violation


@pytest.mark.parametrize("which", [0, 1, 2, 3])
def test_incremental_ddmin_matches_sequential_oracle(world, which):
    eng, dext, recs = world
    i, (ev, par, steps) = list(recs.items())[which]
    rc, mcs_o, st = O.incremental_ddmin(N.MODEL_RAFT5, dext, steps, 2000, O.dpor_seed(ev, par), model_flags=1,
                                        looking_for=1, stop_at_size=1, max_max_distance=64)
    assert rc == 0
    mcs, out = eng.incremental_ddmin(dext, steps, 2000, (ev, par), looking_for=1, stop_at_size=1, max_max_distance=64,
                                     heap_cap=1 << 16)
    assert np.array_equal(mcs, mcs_o)
    assert out.total_replays == st["total_replays"]
    assert out.rounds == st["rounds"]
    assert out.instances == st["instances"]
    assert out.mcs_size == bin(int(mcs_o[0])).count("1")
    assert out.tests_executed >= out.total_replays and out.batches <= out.total_replays


def test_arvind_needs_a_seed(world):
    eng, dext, recs = world
    with pytest.raises(D.DemiError):
        eng.dpor_batch_ex([dext], 10, 10, flags=N.DF_ARVIND_ORDERING)


def test_host_mirror_classes(world):
    """ResumableDPOR / IncrementalDDMin as the reference's drivers use them (RunnerUtils.scala:822-842)."""
    eng, dext, recs = world
    i, (ev, par, steps) = list(recs.items())[0]
    prog = [e for e in D.raft5_program() if e.kind in (N.EXT_START, N.EXT_SEND)]
    cfg = D.SchedulerConfig(N.MODEL_RAFT5, model_flags=1)
    heuristic = D.ArvindDistanceOrdering()
    heuristic.init(ev, par)
    oracle = D.ResumableDPOR(cfg, ev, par, max_messages=steps, max_interleavings=2000, engine=eng,
                             backtrackHeuristic=heuristic)
    ddmin = D.IncrementalDDMin(oracle, maxMaxDistance=64, stopAtSize=1)
    mcs = ddmin.minimize(prog, 1)
    rc, mcs_o, st = O.incremental_ddmin(N.MODEL_RAFT5, D.pack_externals(prog), steps, 2000, O.dpor_seed(ev, par),
                                        model_flags=1, looking_for=1, stop_at_size=1, max_max_distance=64)
    assert [e._id for e in mcs] == [e._id for k, e in enumerate(prog) if (int(mcs_o[0]) >> k) & 1]
    assert ddmin._stats.total_replays == st["total_replays"]
    # the instance of the full sequence: found under cap 0, and answers at once afterwards
    oracle.setMaxDistance(0)
    assert oracle.test(prog, 1) and oracle.test(prog, 1)
