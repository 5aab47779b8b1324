"""Host-side Fuzzer / experiment format."""
import numpy as np

import demi_b200 as D
from demi_b200 import fuzzer as F, experiment as X, _native as N


def test_java_random_next_double_known_values(oracle):
    assert F.JavaRandom(42).nextDouble() == 0.7275636800328681      # Java SE spec value
    assert F.JavaRandom(0).nextDouble() == 0.730967787376657
    for seed, bound in ((42, 0), (42, 10), (7, 16), (12345, (1 << 30) + 1), (-5, 3)):
        r = F.JavaRandom(seed)
        got = [r.nextInt(bound or None) for _ in range(40)]
        assert got == oracle.kat_jrandom(seed, bound, 40)            # same generator as the C oracle


def test_fuzzer_weights_and_sets():
    w = F.FuzzerWeights()
    assert w.getNextEventType(0.0) == "Kill" and w.getNextEventType(0.02) == "Send"
    assert w.getNextEventType(0.55) == "Partition" and w.getNextEventType(0.75) == "UnPartition"
    assert w.getNextEventType(0.99) is None                          # WaitQuiescence
    s = F.RandomizedHashSet(1)
    for v in "abcd":
        s.insert(v)
    s.arr[1] = s.arr[-1]; s.arr.pop()                                # remove index 1 -> [a, d, c] (Util.scala:155-160)
    assert s.arr == ["a", "d", "c"]


def test_generate_fuzz_test_is_well_formed_and_reproducible():
    prefix = [D.Start(a) for a in range(5)] + [D.Send(a, 1, 0x1F) for a in range(5)]
    f = F.Fuzzer(60, F.FuzzerWeights(kill=0.02, send=0.4), F.ClientCommandGenerator(2), prefix, seed=11)
    t1 = f.generateFuzzTest()
    f.message_gen.counter = 0
    t2 = f.generateFuzzTest()
    assert [(type(e), e.a, e.b, e.p0) for e in t1] == [(type(e), e.a, e.b, e.p0) for e in t2]
    assert t1[:10] == prefix and isinstance(t1[-1], D.WaitQuiescence)
    for a, b in zip(t1, t1[1:]):
        assert not (isinstance(a, D.WaitQuiescence) and isinstance(b, D.WaitQuiescence))
    killed, parts = set(), set()
    for e in t1[10:]:
        if isinstance(e, D.Kill):
            assert e.a not in killed
            killed.add(e.a)
        elif isinstance(e, D.Send):
            assert e.a not in killed                                 # only alive actors receive
        elif isinstance(e, D.Partition):
            assert (e.a, e.b) not in parts and e.a < e.b
            parts.add((e.a, e.b))
        elif isinstance(e, D.UnPartition):
            assert (e.a, e.b) in parts
            parts.remove((e.a, e.b))
    other = F.Fuzzer(60, F.FuzzerWeights(kill=0.02, send=0.4), F.ClientCommandGenerator(2), prefix, seed=12).generateFuzzTest()
    assert [(type(e), e.a) for e in other] != [(type(e), e.a) for e in t1]


def test_fuzz_test_runs_on_the_oracle(oracle):
    prefix = [D.Start(a) for a in range(3)]
    f = F.Fuzzer(40, F.FuzzerWeights(send=0.5), F.ClientCommandGenerator(1), prefix, seed=3)
    prog = f.generateFuzzTest()
    ext = D.pack_externals(prog)
    res = oracle.fuzz_batch(N.MODEL_PINGPONG3, ext, 1, 200, -1, 0)
    assert (res["status"] == 0).all() and (res["steps"] > 0).all()


def test_experiment_roundtrip(tmp_path, oracle):
    ext = D.pack_externals(D.raft5_program(client_cmds=2))
    ev, par, r = oracle.fuzz_trace(N.MODEL_RAFT5, ext, 7, 50, 5, model_flags=1)
    mcs = oracle.full_mask(ext)
    X.save_experiment(str(tmp_path / "exp"), N.MODEL_RAFT5, 1, ext, ev, int(r["violation"]), par, mcs,
                      seed=7, max_messages=50, total_replays=3)
    e = X.load_experiment(str(tmp_path / "exp"))
    assert (e["externals"] == ext).all() and (e["events"] == ev).all() and (e["dep_parent"] == par).all()
    assert (e["mcs"] == mcs).all() and e["meta"]["seed"] == 7 and e["meta"]["model"] == N.MODEL_RAFT5
    assert (tmp_path / "exp" / "event_trace.bin").stat().st_size == 16 * len(ev)
