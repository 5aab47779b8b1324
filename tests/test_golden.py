"""Committed fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle).
CPU: the oracle still reproduces them.  GPU: the engine reproduces them through the C ABI."""
import glob
import os

import numpy as np
import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NOT_BATCHES = ("raft5_trace_ddmin.npz", "raft5_round1_extras.npz")
BATCHES = sorted(p for p in glob.glob(os.path.join(HERE, "*.npz")) if not p.endswith(NOT_BATCHES))


@pytest.mark.parametrize("path", BATCHES, ids=[os.path.basename(p) for p in BATCHES])
def test_oracle_reproduces_golden_batches(path, oracle):
    g = np.load(path)
    model, flags, seed, n, maxm, interval = (int(x) for x in g["params"])
    res = oracle.fuzz_batch(model, g["ext"], seed, n, maxm, interval, model_flags=flags)
    assert (res == g["results"]).all()


def test_oracle_reproduces_golden_trace_and_ddmin(oracle):
    g = np.load(os.path.join(HERE, "raft5_trace_ddmin.npz"))
    ev, par, r = oracle.fuzz_trace(2, g["ext"], int(g["seed"]), 50, 5, model_flags=1)
    assert (ev == g["events"]).all() and (par == g["dep_parent"]).all() and r == g["result"][0]
    rc, mcs, total, iters, ver = oracle.ddmin_sts(2, ev, g["ext"], int(r["violation"]), model_flags=1)
    assert rc == 0 and (mcs == g["mcs"]).all() and total == int(g["total_replays"]) and ver == int(g["verified"])
    assert (iters == g["iteration_sizes"]).all()
    rep = oracle.replay_batch(2, ev, g["ext"], g["masks"], looking_for=int(r["violation"]), model_flags=1)
    assert (rep == g["replay"]).all()


def test_oracle_reproduces_round1_extras(oracle):
    """provenance, plain / seeded / capped DPOR, IncrementalDDMin and both internal-minimization strategies on the
    golden execution (tests/golden/raft5_round1_extras.npz)."""
    g = np.load(os.path.join(HERE, "raft5_trace_ddmin.npz"))
    x = np.load(os.path.join(HERE, "raft5_round1_extras.npz"))
    ext, ev, par = g["ext"], g["events"], g["dep_parent"]
    seed, code, steps = int(g["seed"]), int(g["result"][0]["violation"]), int(g["result"][0]["steps"])
    keep, po = oracle.fuzz_provenance(2, ext, seed, 50, 5, 1, model_flags=1)
    assert (keep == x["prov_keep"]).all() and po == x["prov_out"][0]
    dext = ext[(ext["kind"] == 1) | (ext["kind"] == 3)]
    rc, dr, dv, dh = oracle.dpor_search(2, dext, 40, 60, model_flags=1, node_cap=4096, explored_slots=1 << 16, heap_cap=1 << 16)
    assert rc == 0 and dr == x["dpor_result"][0] and (dh == x["dpor_hashes"]).all()
    sd = oracle.dpor_seed(ev, par)
    rows, hashes = [], []
    for arvind, prio, caps in ((1, 1, [0, 2, 4, 8, -1]), (0, 1, [0, 0, -1]), (0, 0, [-1])):
        inst = oracle.DporInstance(2, dext, steps, 80, seed=sd, arvind=arvind, prioritize_pending=prio, model_flags=1,
                                   looking_for=code)
        for c in caps:
            rr, hh = inst.test(c)
            rows.append(rr); hashes.append(hh)
        inst.close()
    assert (np.array(rows) == x["inst_results"]).all() and (np.concatenate(hashes) == x["inst_hashes"]).all()
    rc, mcs, st = oracle.incremental_ddmin(2, dext, steps, 2000, sd, model_flags=1, looking_for=code, stop_at_size=1,
                                           max_max_distance=64)
    assert rc == 0 and (mcs == x["inc_mcs"]).all()
    assert [st["total_replays"], st["rounds"], st["interleavings"], st["instances"]] == x["inc_stats"].tolist()
    rr, vtrace = oracle.replay_trace(2, ev, ext, oracle.full_mask(ext), looking_for=code, model_flags=1)
    mext = ext[ext["kind"] != 4]
    for name, fl in (("ltr", 0), ("fifo", 0x100)):
        rc, tr, total, sizes, unig = oracle.internal_minimize(2, vtrace, mext, code, model_flags=1, flags=fl)
        assert [rc, total, unig] == x["im_%s_meta" % name].tolist()
        assert len(tr) == len(x["im_%s_trace" % name]) and (tr == x["im_%s_trace" % name]).all()
        assert (sizes == x["im_%s_sizes" % name]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("path", BATCHES, ids=[os.path.basename(p) for p in BATCHES])
def test_engine_reproduces_golden_batches(path):
    import demi_b200 as D
    g = np.load(path)
    model, flags, seed, n, maxm, interval = (int(x) for x in g["params"])
    eng = D.Engine(D.SchedulerConfig(model, model_flags=flags))
    eng.set_externals(g["ext"])
    assert (eng.fuzz_batch(seed, n, maxm, interval) == g["results"]).all()


@pytest.mark.gpu
def test_engine_reproduces_golden_trace_and_ddmin():
    import demi_b200 as D
    g = np.load(os.path.join(HERE, "raft5_trace_ddmin.npz"))
    eng = D.Engine(D.SchedulerConfig(2, model_flags=1))
    eng.set_externals(g["ext"])
    ev, par, r = eng.fuzz_trace(int(g["seed"]), 50, 5)
    assert (ev == g["events"]).all() and (par == g["dep_parent"]).all() and r == g["result"][0]
    eng.set_trace(ev, g["ext"])
    code = int(r["violation"])
    assert (eng.replay_batch(g["masks"], code) == g["replay"]).all()
    mcs, iters, out = eng.ddmin(code)
    assert (mcs == g["mcs"]).all() and out.total_replays == int(g["total_replays"]) and out.verified == int(g["verified"])
    assert (iters == g["iteration_sizes"]).all()
