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


def test_c_abi_fuzzer_equals_the_python_mirror(native):
    """demi_fuzzer_generate (the C ABI's Fuzzer.generateFuzzTest) draws the same program as the Python restatement."""
    import ctypes as C
    L = native.lib()
    prefix = [D.Start(a) for a in range(5)] + [D.Send(a, 1, 0x1F) for a in range(5)]
    pext = D.pack_externals(prefix)
    for seed, kw in [(11, dict(kill=0.02, send=0.4)), (5, dict()), (-3, dict(kill=0.2, send=0.1, partition=0.3, unpartition=0.3)),
                     (99, dict(kill=0.5, send=0.1))]:
        w = F.FuzzerWeights(**kw)
        py = F.Fuzzer(80, w, F.ClientCommandGenerator(2), prefix, seed=seed).generateFuzzTest()
        full = dict(kill=0.01, send=0.3, wait_quiescence=0.1, partition=0.1, unpartition=0.1); full.update(kw)
        cfg = native.FuzzerConfig(full["kill"], full["send"], full["wait_quiescence"], full["partition"], full["unpartition"], 80, 2)
        out = np.zeros(256, dtype=native.EXT_DTYPE)
        n = C.c_uint32()
        rc = L.demi_fuzzer_generate(C.byref(cfg), seed, pext.ctypes.data, len(pext), None, 0, out.ctypes.data, len(out), C.byref(n))
        assert rc == 0, L.demi_last_error(None)
        got = out[:n.value]
        want = D.pack_externals(py)
        assert len(got) == len(want)
        for f in ("kind", "a", "b", "type", "p0", "p1"):
            assert (got[f] == want[f]).all(), (seed, f)
        assert len(set(int(x) for x in got["id"])) == len(got)            # UniqueExternalEvent ids stay unique


def test_c_abi_experiment_directory_round_trip(native, tmp_path):
    import ctypes as C
    L = native.lib()
    ext = D.pack_externals(D.raft5_program(client_cmds=3))
    ev = np.zeros(7, dtype=native.EVENT_DTYPE); ev["kind"] = 1; ev["uniq"] = np.arange(7)
    par = np.arange(9, dtype=np.uint16); mcs = np.array([0x5A5], dtype=np.uint64)
    e = native.Experiment(2, 1, 1, 0, ext.ctypes.data, len(ext), len(ext), ev.ctypes.data, len(ev), len(ev),
                          par.ctypes.data, len(par), len(par), mcs.ctypes.data, 1, 1)
    d = str(tmp_path / "exp").encode()
    assert L.demi_experiment_save(d, C.byref(e)) == 0
    back = X.load_experiment(d.decode())                                     # the Python reader sees the same files
    assert (back["externals"] == ext).all() and (back["events"] == ev).all() and (back["dep_parent"] == par).all()
    assert back["meta"]["model"] == 2 and back["meta"]["violation"] == 1 and int(back["mcs"][0]) == 0x5A5
    probe = native.Experiment()
    assert L.demi_experiment_load(d, C.byref(probe)) == native.ERR_CAPACITY  # sizes first
    assert (probe.n_externals, probe.n_events, probe.n_nodes, probe.mask_words) == (len(ext), 7, 9, 1)
    ext2 = np.zeros(probe.n_externals, dtype=native.EXT_DTYPE); ev2 = np.zeros(probe.n_events, dtype=native.EVENT_DTYPE)
    par2 = np.zeros(probe.n_nodes, dtype=np.uint16); mcs2 = np.zeros(1, dtype=np.uint64)
    e2 = native.Experiment(0, 0, 0, 0, ext2.ctypes.data, 0, len(ext2), ev2.ctypes.data, 0, len(ev2), par2.ctypes.data, 0, len(par2),
                           mcs2.ctypes.data, 0, 1)
    assert L.demi_experiment_load(d, C.byref(e2)) == 0
    assert (ext2 == ext).all() and (ev2 == ev).all() and (par2 == par).all() and mcs2[0] == mcs[0]
    assert (e2.model, e2.model_flags, e2.violation) == (2, 1, 1)
    X.save_experiment(str(tmp_path / "exp2"), 2, 1, ext, ev, 1, dep_parent=par, mcs=mcs)   # and the C reader the Python writer's
    e3 = native.Experiment(0, 0, 0, 0, ext2.ctypes.data, 0, len(ext2), ev2.ctypes.data, 0, len(ev2), None, 0, 0, None, 0, 0)
    assert L.demi_experiment_load(str(tmp_path / "exp2").encode(), C.byref(e3)) == 0 and e3.violation == 1
