"""Golden traces recorded by the REAL DEMi on a JVM (oracle/jvm/Runner) against the CPU oracle.

The image has no JVM, so tests/golden/jvm/ is empty and this test skips; the day a dump exists it pins the oracle's
restatement of RandomScheduler / DepTracker to the reference event by event (ids normalised per execution)."""
import glob
import json
import os

import numpy as np
import pytest

import demi_b200 as D
from demi_b200 import experiment as X

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jvm")


def dumps():
    return sorted(d for d in glob.glob(os.path.join(GOLDEN, "*")) if os.path.exists(os.path.join(d, "meta.json")))


@pytest.mark.skipif(not dumps(), reason="no JVM dumps under tests/golden/jvm (this image has no JDK; see oracle/jvm/README.md)")
def test_oracle_reproduces_the_jvm_traces(oracle):
    for d in dumps():
        e = X.load_experiment(d)
        name = os.path.basename(d)                                   # <model>_seed<seed>[_m<maxMessages>_i<interval>]
        seed = int(name.split("seed")[1].split("_")[0])
        maxm = int(name.split("_m")[1].split("_")[0]) if "_m" in name else (50 if e["meta"]["model"] == 2 else -1)
        interval = int(name.split("_i")[1].split("_")[0]) if "_i" in name else (5 if e["meta"]["model"] == 2 else 0)
        ev, par, r = oracle.fuzz_trace(e["meta"]["model"], e["externals"], seed, maxm, interval, model_flags=e["meta"]["model_flags"])
        assert int(r["violation"]) == int(e["meta"]["violation"]), d
        assert len(ev) == len(e["events"]), d
        for f in ("kind", "src", "dst", "type", "p0", "p1", "uniq"):   # node ids are not recorded by the JVM dump
            assert (ev[f] == e["events"][f]).all(), (d, f)
