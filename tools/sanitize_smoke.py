"""Small invocation of every kernel, meant to run under compute-sanitizer (memcheck / racecheck)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import demi_b200 as D
from demi_b200 import _native as N
from oracle import binding as O

R2_ONLY = bool(os.environ.get("SMOKE_R2_ONLY"))
for strategy in (() if R2_ONLY else (0, 1)):
    for model, prog, flags, maxm, iv in ((2, D.raft5_program(client_cmds=2), 1, 50, 5), (1, D.pingpong3_program(20), 0, -1, 0),
                                         (3, D.bcast32_program(2), 0, 60, 0)):
        ext = D.pack_externals(prog)
        for lane in (True, False):
            if lane:
                os.environ.pop("DEMI_DISABLE_LANE_ENGINE", None)
            else:
                os.environ["DEMI_DISABLE_LANE_ENGINE"] = "1"
            eng = D.Engine(D.SchedulerConfig(model, model_flags=flags, strategy=strategy))
            eng.set_externals(ext)
            g = eng.fuzz_batch(1, 300, maxm, iv, flags=1)
            c = O.fuzz_batch(model, ext, 1, 300, maxm, iv, model_flags=flags, flags=1, strategy=strategy)
            assert (g == c).all(), (model, strategy, lane)
            eng.fuzz_trace(3, maxm, iv)
            eng.close()
os.environ.pop("DEMI_DISABLE_LANE_ENGINE", None)
prog = D.raft5_program(client_cmds=4)
ext = D.pack_externals(prog)
res = O.fuzz_batch(2, ext, 1, 3000, 50, 5, model_flags=1)
seed = 1 + int(np.nonzero(res["violation"])[0][0])
ev, par, r = O.fuzz_trace(2, ext, seed, 50, 5, model_flags=1)
code = int(r["violation"])
eng = D.Engine(D.SchedulerConfig(2, model_flags=1))
eng.set_trace(ev, ext)
if not R2_ONLY:
    rng = np.random.default_rng(0)
    masks = (rng.integers(0, 2 ** len(ext), size=400, dtype=np.uint64) & O.full_mask(ext)[0]).reshape(-1, 1)
    for fl in (0, 1, 2):
        assert (eng.replay_batch(masks, code, fl) == O.replay_batch(2, ev, ext, masks, looking_for=code, flags=fl, model_flags=1)).all()
    mcs, iters, dd = eng.ddmin(code)
    rr, vtrace = eng.replay_trace(O.full_mask(ext), looking_for=code)
    mext = D.pack_externals([e for e in prog if not isinstance(e, D.WaitQuiescence)])
    eng.set_trace(vtrace, mext)
    eng.internal_minimize(code)
    progs = [[D.Start(a) for a in range(5)] + [D.Send(a, 1, 0x1F) for a in range(k)] for k in (2, 3, 5)] * 4
    eng.dpor_batch(progs, 30, 40)
    u, i = eng.dedup_compact(eng.fuzz_batch(1, 5000, 6, 5, flags=1) if eng.set_externals(ext) is None else None, 0)
    # provenance pruning, seeded / capped DPOR instances, IncrementalDDMin
    res = eng.fuzz_batch(1, 2000, 40, 5)
    viol = np.nonzero(res["violation"] == 1)[0].astype(np.uint32)
    keep, pout, _ = eng.fuzz_provenance(1, viol[:40], 40, 5)
    assert (pout["status"] == 0).all()
    sev, spar, sr = eng.fuzz_trace(1 + int(viol[0]), 40, 5)
    k2, o2 = eng.provenance(sev, spar, int(pout[0]["affected_mask"]))
    assert (k2[:len(keep[0])] == keep[0][:len(k2)]).all()
    dext = ext[(ext["kind"] == 1) | (ext["kind"] == 3)]
    steps = int(sr["steps"])
    caps = [[0, 2, 4, -1], [0], [0, 2], [-1]]
    progs2 = [dext, dext[1:], dext[:-2], dext]
    for fl in (0, 3):
        r2, h2 = eng.dpor_batch_ex(progs2, steps, 40, seed=(sev, spar), flags=fl, caps=caps, looking_for=1, heap_cap=1 << 15,
                                   want_hashes=True)
        assert (r2["status"] == 0).all()
    m2, io = eng.incremental_ddmin(dext, steps, 200, (sev, spar), looking_for=1, stop_at_size=1, max_max_distance=8)
else:
    u, io = [], type("o", (), {"mcs_size": 0})()
# ---- round 2: frontier DPOR (history on / off), model IR, user filter + HardKill, lane recording with deferred slots,
# conjoined atoms
from demi_b200 import model_ir
fprog = [D.Start(a) for a in range(5)] + [D.Send(a, 1, 0x1F) for a in range(5)]
for fl, budget in ((0, 4000), (N.FR_NO_HISTORY, 600)):
    F = D.Engine.frontier_params(34, budget, 128, explored_slots=1 << 12, pool_cap=1 << 18, flags=fl)
    fr, fv, fh = eng.dpor_frontier(fprog, F)
    assert int(fr["status"]) == 0, fr
ir = D.Engine(D.SchedulerConfig(N.MODEL_IR, model_flags=1 | (4 << 8)))
ir.load_model(model_ir.pingpong3_blob())
pext = D.pack_externals(D.pingpong3_program(20))
ir.set_externals(pext)
O.load_model(model_ir.pingpong3_blob())
assert (ir.fuzz_batch(1, 300, -1, 3) == O.fuzz_batch(N.MODEL_IR, pext, 1, 300, -1, 3, model_flags=1 | (4 << 8))).all()
ir.close()
hk = D.raft5_program(client_cmds=3)[:-1] + [D.WaitQuiescence(), D.HardKill(1), D.Send(2, 2, 41), D.WaitQuiescence(), D.Start(1),
                                            D.Send(1, 1, 0x1F), D.WaitQuiescence()]
fe = D.Engine(D.SchedulerConfig(2, model_flags=1))
fe.set_user_filter([(0b00110, 0b11111, 1 << 5, 0)])
fe.set_externals(D.pack_externals(hk))
fe.fuzz_batch(1, 300, 90, 7)
fe.close()
os.environ["DEMI_LANE_PENDING_CAP"] = "24"
le = D.Engine(D.SchedulerConfig(2, model_flags=1))
le.set_externals(D.pack_externals(D.raft5_program()))
lres = le.fuzz_batch(1, 3000, 50, 5)
lviol = np.nonzero(lres["violation"])[0].astype(np.uint32)
lk, lo, _ = le.fuzz_provenance(1, lviol[:60], 50, 5)
assert (lo["status"] == 0).all() and le.stats().deferred >= 0
le.close()
os.environ.pop("DEMI_LANE_PENDING_CAP")
eng.set_trace(ev, ext)
starts = [i for i in range(len(ext)) if ext[i]["kind"] == 3][:2]
if len(starts) == 2:
    eng.conjoin_atoms(starts[0], starts[1])
    eng.ddmin(code)
print("sanitize smoke ok", len(u), int(io.mcs_size))
