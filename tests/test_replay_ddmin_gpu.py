"""GPU parity for K2: batched STSSched replay and DDMin vs the CPU oracle, bit-exact."""
import numpy as np
import pytest

import demi_b200 as D
from demi_b200 import _native as N

pytestmark = pytest.mark.gpu


def violating_trace(oracle, model, prog, flags, maxm, interval, which=0, seeds=20000):
    ext = D.pack_externals(prog)
    res = oracle.fuzz_batch(model, ext, 1, seeds, maxm, interval, model_flags=flags)
    hits = np.nonzero(res["violation"])[0]
    assert len(hits) > which
    seed = 1 + int(hits[which])
    ev, par, r = oracle.fuzz_trace(model, ext, seed, maxm, interval, model_flags=flags)
    return ext, ev, int(r["violation"])


def random_masks(ext, n, seed, oracle):
    rng = np.random.default_rng(seed)
    mw = oracle.mask_words(len(ext))
    full = oracle.full_mask(ext)
    m = rng.integers(0, 2**63, size=(n, mw), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, mw), dtype=np.uint64)
    dens = rng.integers(0, 3, size=n)
    extra = rng.integers(0, 2**63, size=(n, mw), dtype=np.uint64) * np.uint64(2) + np.uint64(1)
    m = np.where((dens == 0)[:, None], m, m | extra)       # mix of ~50% and ~75% dense masks
    m = np.where((dens == 2)[:, None], m | ~extra, m)
    m &= full[None, :]
    m[0] = full
    m[1] = 0
    return m


@pytest.mark.parametrize("fka", [0, 1])
def test_replay_batch_matches_oracle_raft(oracle, fka):
    prog = D.raft5_program(client_cmds=6)
    ext, ev, code = violating_trace(oracle, N.MODEL_RAFT5, prog, 1, 50, 5)
    eng = D.Engine(D.SchedulerConfig(N.MODEL_RAFT5, model_flags=1))
    eng.set_trace(ev, ext)
    masks = random_masks(ext, 3000, 1, oracle)
    flags = N.RF_FILTER_KNOWN_ABSENTS if fka else 0
    gpu = eng.replay_batch(masks, code, flags)
    cpu = oracle.replay_batch(N.MODEL_RAFT5, ev, ext, masks, looking_for=code, flags=flags, model_flags=1)
    bad = np.nonzero(gpu != cpu)[0]
    assert len(bad) == 0, (bad[:5], gpu[bad[:3]], cpu[bad[:3]])
    assert gpu[0]["violation"] == code and gpu[0]["ignored"] == 0     # the full subsequence reproduces it
    assert gpu[1]["violation"] == 0 and gpu[1]["delivered"] == 0      # the empty one delivers nothing
    assert (gpu["status"] == 0).all()


def test_replay_with_kills_partitions_pingpong(oracle):
    ev_prog = [D.Start(a) for a in range(3)]
    ev_prog += [D.Send(k % 3, 1, k) for k in range(10)]
    ev_prog += [D.WaitQuiescence(), D.Partition(0, 1), D.Kill(2)]
    ev_prog += [D.Send(k % 3, 1, 100 + k) for k in range(10)]
    ev_prog += [D.WaitQuiescence(), D.UnPartition(0, 1), D.Start(2)]
    ev_prog += [D.Send(k % 3, 1, 200 + k) for k in range(9)]
    ev_prog += [D.WaitQuiescence()]
    ext = D.pack_externals(ev_prog)
    flags = 1 | (6 << 8)         # PingPong3 test hook: violation 7 once actor 0 has >= 6 pongs
    ev, par, r = oracle.fuzz_trace(N.MODEL_PINGPONG3, ext, 5, -1, 0, model_flags=flags)
    assert r["violation"] == 7
    eng = D.Engine(D.SchedulerConfig(N.MODEL_PINGPONG3, model_flags=flags))
    eng.set_trace(ev, ext)
    masks = random_masks(ext, 4000, 2, oracle)
    for fl in (0, N.RF_FILTER_KNOWN_ABSENTS):
        gpu = eng.replay_batch(masks, 7, fl)
        cpu = oracle.replay_batch(N.MODEL_PINGPONG3, ev, ext, masks, looking_for=7, flags=fl, model_flags=flags)
        assert (gpu == cpu).all()
    assert (gpu["violation"] != 0).sum() > 0 and (gpu["violation"] == 0).sum() > 0


def test_replay_bcast32(oracle):
    prog = D.bcast32_program(2)
    ext = D.pack_externals(prog)
    ev, par, r = oracle.fuzz_trace(N.MODEL_BCAST32, ext, 3, 120, 10, model_flags=3)
    assert r["violation"] == 3
    eng = D.Engine(D.SchedulerConfig(N.MODEL_BCAST32, model_flags=3))
    eng.set_trace(ev, ext)
    masks = random_masks(ext, 500, 3, oracle)
    gpu = eng.replay_batch(masks, 3)
    cpu = oracle.replay_batch(N.MODEL_BCAST32, ev, ext, masks, looking_for=3, model_flags=3)
    assert (gpu == cpu).all()


def test_strict_replay_validates_and_diverges(oracle):
    """ReplayScheduler (ReplayScheduler.scala:256-342): the recorded trace replays strictly; a trace of
    another seed diverges (ReplayException)."""
    prog = D.raft5_program(client_cmds=2)
    ext, ev, code = violating_trace(oracle, N.MODEL_RAFT5, prog, 1, 50, 5)
    cfg = D.SchedulerConfig(N.MODEL_RAFT5, model_flags=1)
    rs = D.ReplayScheduler(cfg, ev, prog)
    r = rs.replay(code)
    assert r["violation"] == code and r["ignored"] == 0
    # drop one delivery's send from the trace: its delivery can no longer be pending
    first_delivery = np.nonzero(ev["kind"] == N.EV_MSG_EVENT)[0][3]
    broken = ev.copy()
    broken[first_delivery]["p0"] ^= 0x55
    rs2 = D.ReplayScheduler(cfg, broken, prog)
    with pytest.raises(D.ReplayScheduler.ReplayException):
        rs2.replay(code)
    cpu = oracle.replay_batch(N.MODEL_RAFT5, broken, ext, [oracle.full_mask(ext)], looking_for=code, flags=2, model_flags=1)
    assert cpu[0]["status"] == N.RS_DIVERGED


@pytest.mark.parametrize("which", [0, 1, 2])
def test_ddmin_matches_sequential_oracle(oracle, which):
    prog = D.raft5_program(client_cmds=14)
    ext, ev, code = violating_trace(oracle, N.MODEL_RAFT5, prog, 1, 50, 5, which=which)
    cfg = D.SchedulerConfig(N.MODEL_RAFT5, model_flags=1)
    sts = D.STSScheduler(cfg, ev, prog)
    dd = D.DDMin(sts, checkUnmodifed=True)
    mcs = dd.minimize(code)
    rc, cmcs, total_replays, iters, verified = oracle.ddmin_sts(N.MODEL_RAFT5, ev, ext, code, model_flags=1)
    assert rc == 0
    assert (D.mask_of(prog, mcs) == cmcs).all()
    assert dd._stats.total_replays == total_replays                  # sequential-equivalent stats
    assert dd._stats.iteration_size == [int(x) for x in iters]
    assert dd.last.verified == verified == 1
    assert dd.last.replays_executed >= total_replays                 # speculation only ever adds
    assert dd.verify_mcs(mcs, code) is not None                      # verify_mcs (DeltaDebugging.scala:64-71)
    assert len(mcs) < len([e for e in prog if not isinstance(e, D.WaitQuiescence)])


def test_ddmin_with_atomic_pairs_pingpong(oracle):
    """Start..Kill and Partition..UnPartition are removed atomically (minification/Util.scala:197-265)."""
    prog = [D.Start(a) for a in range(3)]
    prog += [D.Send(k % 3, 1, k) for k in range(16)]
    prog += [D.WaitQuiescence(), D.Partition(1, 2), D.Kill(2)]
    prog += [D.Send(k % 2, 1, 100 + k) for k in range(16)]
    prog += [D.WaitQuiescence(), D.UnPartition(1, 2)]
    prog += [D.Send(k % 2, 1, 200 + k) for k in range(8)]
    prog += [D.WaitQuiescence()]
    ext = D.pack_externals(prog)
    flags = 1 | (5 << 8)
    ev, par, r = oracle.fuzz_trace(N.MODEL_PINGPONG3, ext, 11, -1, 0, model_flags=flags)
    assert r["violation"] == 7
    cfg = D.SchedulerConfig(N.MODEL_PINGPONG3, model_flags=flags)
    sts = D.STSScheduler(cfg, ev, prog)
    dd = D.DDMin(sts, checkUnmodifed=True)
    mcs = dd.minimize(7)
    rc, cmcs, total_replays, iters, verified = oracle.ddmin_sts(N.MODEL_PINGPONG3, ev, ext, 7, model_flags=flags)
    assert rc == 0 and (D.mask_of(prog, mcs) == cmcs).all()
    assert dd._stats.total_replays == total_replays and dd._stats.iteration_size == [int(x) for x in iters]
    assert dd.last.verified == verified
    # passing oracle for a fingerprint that never occurs: "Unmodified trace does not trigger violation"
    with pytest.raises(D.DemiError):
        D.DDMin(sts, checkUnmodifed=True).minimize(5)


def test_replay_large_batch_properties(oracle):
    """BASELINE config[3]-sized batch: idempotence/monotone properties that need no oracle."""
    prog = D.raft5_program(client_cmds=30)
    ext, ev, code = violating_trace(oracle, N.MODEL_RAFT5, prog, 1, 50, 5)
    eng = D.Engine(D.SchedulerConfig(N.MODEL_RAFT5, model_flags=1))
    eng.set_trace(ev, ext)
    masks = random_masks(ext, 200_000, 9, oracle)
    a = eng.replay_batch(masks, code)
    b = eng.replay_batch(masks, code)
    assert (a == b).all() and (a["status"] == 0).all()               # deterministic
    n_deliveries = int((ev["kind"] == N.EV_MSG_EVENT).sum())
    assert ((a["delivered"].astype(int) + a["ignored"]) <= n_deliveries).all()
    assert a[0]["violation"] == code
    # equal masks give equal results wherever they sit in the batch
    perm = np.random.default_rng(1).permutation(len(masks))
    c = eng.replay_batch(masks[perm], code)
    assert (c == a[perm]).all()
    st = eng.stats()
    assert st.violations == int((c["violation"] != 0).sum())


def test_ddmin_with_conjoined_atoms(oracle):
    """UnmodifiedEventDag.conjoinAtoms (minification/Util.scala:167-178): two externals the caller ties together are
    kept or removed as one atom; MCS and sequential stats still equal the oracle's."""
    prog = D.raft5_program(client_cmds=14)
    ext, ev, code = violating_trace(oracle, N.MODEL_RAFT5, prog, 1, 50, 5, which=1)
    cfg = D.SchedulerConfig(N.MODEL_RAFT5, model_flags=1)
    sts = D.STSScheduler(cfg, ev, prog)
    dd = D.DDMin(sts, checkUnmodifed=True)
    pairs = [(11, 17), (12, 20)]
    for i, j in pairs:
        dd.conjoinAtoms(prog[i], prog[j])
    with pytest.raises(D.DemiError):
        dd.conjoinAtoms(prog[11], prog[13])                          # the assert at :174: already conjoined
    mcs = dd.minimize(code)
    oracle.set_conjoined(pairs, len(ext))
    try:
        rc, cmcs, total_replays, iters, verified = oracle.ddmin_sts(N.MODEL_RAFT5, ev, ext, code, model_flags=1)
    finally:
        oracle.set_conjoined([], len(ext))
    assert rc == 0 and (D.mask_of(prog, mcs) == cmcs).all()
    assert dd._stats.total_replays == total_replays and dd._stats.iteration_size == [int(x) for x in iters]
    m = int(cmcs[0])
    for i, j in pairs:
        assert ((m >> i) & 1) == ((m >> j) & 1)
    json_line = dd._stats.toJson()
    assert '"iteration_size"' in json_line and '"total_replays": %d' % total_replays in json_line
