"""The model IR (include/demi_model_ir.h): pingpong3 re-expressed as a demi_load_model blob behaves exactly like the
compiled model — in the CPU interpreter here, on the device in the gpu tests."""
import numpy as np
import pytest

import demi_b200 as D
from demi_b200 import model_ir, _native as N

FIELDS = ["violation", "steps", "trace_hash", "n_nodes", "n_events", "max_pending", "status"]   # state_hash covers the state geometry


def test_interpreted_pingpong_equals_the_compiled_model_on_the_cpu(oracle):
    oracle.load_model(model_ir.pingpong3_blob())
    ext = D.pack_externals(D.pingpong3_program(20))
    flags = 1 | (4 << 8)
    a = oracle.fuzz_batch(N.MODEL_PINGPONG3, ext, 1, 3000, -1, 3, model_flags=flags)
    b = oracle.fuzz_batch(N.MODEL_IR, ext, 1, 3000, -1, 3, model_flags=flags)
    for f in FIELDS:
        assert (a[f] == b[f]).all(), f
    assert (a["violation"] == 7).any()
    for seed in (1, 5, 77):
        ev1, par1, _ = oracle.fuzz_trace(N.MODEL_PINGPONG3, ext, seed, -1, 3, model_flags=flags)
        ev2, par2, _ = oracle.fuzz_trace(N.MODEL_IR, ext, seed, -1, 3, model_flags=flags)
        assert (ev1 == ev2).all() and (par1 == par2).all()


def test_assembler_resolves_labels_and_encodes_operands():
    w = model_ir.assemble([("LDI", "r3", 9), ("label", "top"), ("ADD", "r1", "r2", "r3"), ("JNE", "r1", "r3", "top"), ("HALT",)])
    assert w == [1 | (3 << 8), 9, 3 | (1 << 8) | (2 << 16) | (3 << 24), 17 | (1 << 8) | (3 << 16), 2, 0]


@pytest.mark.gpu
def test_loaded_model_on_the_device(oracle):
    blob = model_ir.pingpong3_blob()
    oracle.load_model(blob)
    prog = D.pingpong3_program(20)
    ext = D.pack_externals(prog)
    flags = 1 | (4 << 8)
    eng = D.Engine(D.SchedulerConfig(N.MODEL_IR, model_flags=flags))
    with pytest.raises(D.DemiError) as ei:
        eng.set_externals(ext)                                       # no model yet
    assert ei.value.code == N.ERR_STATE
    with pytest.raises(D.DemiError):
        eng.load_model(blob[:40])                                    # truncated
    bad = bytearray(blob); bad[16 * 4] = 0x7F                        # unknown opcode in receive()
    with pytest.raises(D.DemiError):
        eng.load_model(bytes(bad))
    eng.load_model(blob)
    assert eng.actor_index("B") == 1 and eng.actor_name(2) == "C" and eng.actor_index("nobody") == -1
    eng.set_externals(ext)
    n = 20000
    gpu = eng.fuzz_batch(1, n, -1, 3)
    cpu = oracle.fuzz_batch(N.MODEL_IR, ext, 1, n, -1, 3, model_flags=flags)
    assert (gpu == cpu).all()                                        # every field, state hash included
    comp = D.Engine(D.SchedulerConfig(N.MODEL_PINGPONG3, model_flags=flags))
    comp.set_externals(ext)
    ref = comp.fuzz_batch(1, n, -1, 3)
    for f in FIELDS:
        assert (gpu[f] == ref[f]).all(), f                           # ... and the compiled model agrees
    hits = np.nonzero(gpu["violation"])[0]
    ev, par, r = eng.fuzz_trace(1 + int(hits[0]), -1, 3)
    cev, cpar, _ = oracle.fuzz_trace(N.MODEL_IR, ext, 1 + int(hits[0]), -1, 3, model_flags=flags)
    assert (ev == cev).all() and (par == cpar).all()
    # DDMin over STSSched replays of the loaded model
    eng.set_trace(ev, ext)
    mcs, iters, dd = eng.ddmin(7)
    rc, cmcs, total, citers, ver = oracle.ddmin_sts(N.MODEL_IR, ev, ext, 7, model_flags=flags)
    assert rc == 0 and (mcs == cmcs).all() and dd.total_replays == total and list(iters) == list(citers) and dd.verified == 1
    # one DPORwHeuristics search as a frontier
    dprog = [e for e in prog if not isinstance(e, D.WaitQuiescence)][:3 + 6]
    F = eng.frontier_params(14, 20000, 64, explored_slots=1 << 18, pool_cap=1 << 20)
    r, viol, hashes = eng.dpor_frontier(dprog, F)
    OF = oracle.frontier_params(14, 20000, 64, explored_slots=1 << 18, pool_cap=1 << 20)
    rc, ores, oviol, ohashes = oracle.dpor_frontier(N.MODEL_IR, D.pack_externals(dprog), OF, 1, model_flags=flags)
    assert rc == 0 and r["interleavings"] == ores[0]["interleavings"] and (hashes == ohashes[0]).all() and (viol == oviol[0]).all()
    assert r["interleavings"] > 1 and r["exhausted"] == 1
