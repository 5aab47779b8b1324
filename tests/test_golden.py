"""Committed fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle).
CPU: the oracle still reproduces them.  GPU: the engine reproduces them through the C ABI."""
import glob
import os

import numpy as np
import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BATCHES = sorted(p for p in glob.glob(os.path.join(HERE, "*.npz")) if not p.endswith("raft5_trace_ddmin.npz"))


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
