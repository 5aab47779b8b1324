"""GPU parity: recorded STSSched replay + internal-event minimization (STSSchedMinimizer / LeftToRightOneAtATime)."""
import numpy as np
import pytest

import demi_b200 as D
from demi_b200 import _native as N

pytestmark = pytest.mark.gpu


def pipeline(oracle, model, prog, flags, maxm, interval, which=0, code_hint=None):
    ext = D.pack_externals(prog)
    res = oracle.fuzz_batch(model, ext, 1, 20000, maxm, interval, model_flags=flags)
    hits = np.nonzero(res["violation"])[0]
    seed = 1 + int(hits[which])
    ev, par, r = oracle.fuzz_trace(model, ext, seed, maxm, interval, model_flags=flags)
    return ext, ev, int(r["violation"])


@pytest.mark.parametrize("which", [0, 1, 3])
def test_recorded_replay_and_internal_minimization_match_oracle(oracle, which):
    prog = D.raft5_program(client_cmds=5)
    ext, ev, code = pipeline(oracle, N.MODEL_RAFT5, prog, 1, 50, 5, which)
    cfg = D.SchedulerConfig(N.MODEL_RAFT5, model_flags=1)
    eng = D.Engine(cfg)
    eng.set_trace(ev, ext)
    # DDMin, then verify_mcs in recording mode: the trace STSSched returns for the MCS
    mcs_mask, _, dd = eng.ddmin(code)
    r_gpu, vtrace = eng.replay_trace(mcs_mask, looking_for=code)
    r_cpu, vtrace_cpu = oracle.replay_trace(N.MODEL_RAFT5, ev, ext, mcs_mask, looking_for=code, model_flags=1)
    assert r_gpu == r_cpu and len(vtrace) == len(vtrace_cpu) and (vtrace == vtrace_cpu).all()
    if not r_gpu["violation"]:
        # DDMin's result need not reproduce with a non-monotone oracle ("MCS doesn't reproduce bug...",
        # RunnerUtils.scala:699): carry on with the unminimized externals, as the reference's pipeline does
        mcs_mask = oracle.full_mask(ext)
        r_gpu, vtrace = eng.replay_trace(mcs_mask, looking_for=code)
        r_cpu, vtrace_cpu = oracle.replay_trace(N.MODEL_RAFT5, ev, ext, mcs_mask, looking_for=code, model_flags=1)
        assert r_gpu == r_cpu and (vtrace == vtrace_cpu).all()
    assert r_gpu["violation"] == code
    mcs_events = D.events_of(prog, mcs_mask)
    mcs_ext = D.pack_externals(mcs_events)
    # replaying a recorded trace in full reproduces it exactly (fixed point)
    eng.set_trace(vtrace, mcs_ext)
    r2, v2 = eng.replay_trace(None, looking_for=code)
    assert r2["violation"] == code and r2["ignored"] == 0 and len(v2) == len(vtrace) and (v2 == vtrace).all()
    # skip-one-delivery batch == oracle
    deliveries = np.nonzero(vtrace["kind"] == N.EV_MSG_EVENT)[0].astype(np.uint32)
    gpu = eng.replay_batch_ex(None, deliveries, code)
    for j, idx in enumerate(deliveries[:12]):
        rc, _ = oracle.replay_trace(N.MODEL_RAFT5, vtrace, mcs_ext, oracle.full_mask(mcs_ext, False), looking_for=code,
                                    model_flags=1, skip_event=int(idx))
        assert gpu[j] == rc
    # internal minimization
    sm = D.STSSchedMinimizer(mcs_events, vtrace, code, D.LeftToRightOneAtATime(), cfg, engine=eng)
    stats, mtrace = sm.minimize()
    rc, ctrace, total, sizes, unig = oracle.internal_minimize(N.MODEL_RAFT5, vtrace, mcs_ext, code, model_flags=1)
    assert rc == 0
    assert len(mtrace) == len(ctrace) and (mtrace == ctrace).all()
    assert sm.last.total_replays == total and stats.internal_sizes == [int(x) for x in sizes]
    assert sm.last.unignorable == unig
    assert sm.last.deliveries_after <= sm.last.deliveries_before
    assert sm.last.replays_executed >= total
    # the minimized schedule still reproduces the violation, and nothing more can be dropped
    eng.set_trace(mtrace, mcs_ext)
    r3, _ = eng.replay_trace(None, looking_for=code)
    assert r3["violation"] == code and r3["ignored"] == 0
    # SrcDstFIFORemoval (OneAtATimeRemoval.scala:139-251): only the last delivery of each (src,dst) pair, plus timers
    sf = D.STSSchedMinimizer(mcs_events, vtrace, code, D.SrcDstFIFORemoval(), cfg, engine=eng)
    fstats, ftrace = sf.minimize()
    rc, cftrace, ftotal, fsizes, funig = oracle.internal_minimize(N.MODEL_RAFT5, vtrace, mcs_ext, code, model_flags=1,
                                                                  flags=N.IM_SRC_DST_FIFO)
    assert rc == 0
    assert len(ftrace) == len(cftrace) and (ftrace == cftrace).all()
    assert sf.last.total_replays == ftotal and fstats.internal_sizes == [int(x) for x in fsizes]
    eng.set_trace(ftrace, mcs_ext)
    r4, _ = eng.replay_trace(None, looking_for=code)
    assert r4["violation"] == code


def test_internal_minimization_pingpong(oracle):
    flags = 1 | (3 << 8)
    prog = [D.Start(a) for a in range(3)] + [D.Send(k % 3, 1, k) for k in range(18)] + [D.WaitQuiescence()]
    ext = D.pack_externals(prog)
    ev, par, r = oracle.fuzz_trace(N.MODEL_PINGPONG3, ext, 4, -1, 0, model_flags=flags)
    assert r["violation"] == 7
    cfg = D.SchedulerConfig(N.MODEL_PINGPONG3, model_flags=flags)
    eng = D.Engine(cfg)
    eng.set_trace(ev, ext)
    mcs_mask, _, dd = eng.ddmin(7)
    mcs_events = D.events_of(prog, mcs_mask)
    mcs_ext = D.pack_externals(mcs_events)
    r_gpu, vtrace = eng.replay_trace(mcs_mask, looking_for=7)
    r_cpu, vtrace_cpu = oracle.replay_trace(N.MODEL_PINGPONG3, ev, ext, mcs_mask, looking_for=7, model_flags=flags)
    assert r_gpu == r_cpu and (vtrace == vtrace_cpu).all()
    if not r_gpu["violation"]:
        mcs_mask = oracle.full_mask(ext)
        mcs_events = D.events_of(prog, mcs_mask)
        mcs_ext = D.pack_externals(mcs_events)
        r_gpu, vtrace = eng.replay_trace(mcs_mask, looking_for=7)
        assert r_gpu["violation"] == 7
    sm = D.STSSchedMinimizer(mcs_events, vtrace, 7, D.LeftToRightOneAtATime(), cfg, engine=eng)
    stats, mtrace = sm.minimize()
    rc, ctrace, total, sizes, unig = oracle.internal_minimize(N.MODEL_PINGPONG3, vtrace, mcs_ext, 7, model_flags=flags)
    assert rc == 0 and len(mtrace) == len(ctrace) and (mtrace == ctrace).all()
    assert sm.last.total_replays == total and stats.internal_sizes == [int(x) for x in sizes]
