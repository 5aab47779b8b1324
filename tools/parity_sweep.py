"""Full-size parity sweep: the bench workload's prefixes, record by record, GPU engine vs the C oracle."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import demi_b200 as D
from demi_b200 import _native as N
from oracle import binding as O


def sweep(name, model, prog, model_flags, maxm, interval, total, chunk, strategy=0, flags=0):
    ext = D.pack_externals(prog)
    eng = D.Engine(D.SchedulerConfig(model, model_flags=model_flags, strategy=strategy))
    eng.set_externals(ext)
    bad = 0
    viol = 0
    t0 = time.perf_counter()
    for base in range(1, total + 1, chunk):
        g = eng.fuzz_batch(base, chunk, maxm, interval, flags=flags)
        c = O.fuzz_batch(model, ext, base, chunk, maxm, interval, model_flags=model_flags, flags=flags, strategy=strategy)
        bad += int((g != c).sum())
        viol += int((g["violation"] != 0).sum())
    return {"workload": name, "prefixes": total, "mismatching_records": bad, "violating": viol,
            "seconds": round(time.perf_counter() - t0, 1)}


def sweep_dpor(n_searches, max_interleavings, threads):
    """configs[2] at (a slice of) full size: per-search counters and the whole schedule-hash sequence."""
    from concurrent.futures import ThreadPoolExecutor
    rng = np.random.default_rng(7)
    progs = []
    for _ in range(n_searches):
        ev = [D.Start(int(a)) for a in rng.permutation(5)]
        ev += [D.Send(int(a), 1, 0x1F) for a in rng.permutation(5)[:int(rng.integers(3, 6))]]
        ev += [D.Send(int(rng.integers(0, 5)), 2, int(rng.integers(1, 50))) for _ in range(int(rng.integers(0, 3)))]
        progs.append(ev)
    eng = D.Engine(D.SchedulerConfig(N.MODEL_RAFT5, model_flags=3))
    t0 = time.perf_counter()
    res, viol, hashes = eng.dpor_batch(progs, 100, max_interleavings, heap_cap=1 << 17, want_hashes=True)

    def one(i):
        rc, r, v, h = O.dpor_search(N.MODEL_RAFT5, D.pack_externals(progs[i]), 100, max_interleavings, model_flags=3,
                                    node_cap=4096, explored_slots=1 << 16, heap_cap=1 << 17)
        same = all(int(r[f]) == int(res[i][f]) for f in ("interleavings", "violations", "deliveries", "races", "n_nodes",
                                                          "n_explored", "heap_left", "exhausted", "budget_exhausted", "status"))
        return same and hashes[i][:len(h)].tolist() == h.tolist()

    with ThreadPoolExecutor(threads) as ex:
        ok = list(ex.map(one, range(n_searches)))
    return {"workload": "raft5 DPORwHeuristics depth-100, <=%d interleavings per search" % max_interleavings,
            "searches": n_searches, "interleavings": int(res["interleavings"].sum()), "mismatching_searches": int(n_searches - sum(ok)),
            "seconds": round(time.perf_counter() - t0, 1)}


def sweep_replay(n_masks):
    """configs[3]: STSSched replays of random subsequences of a long violating trace."""
    prog = D.raft5_program(client_cmds=290)
    ext = D.pack_externals(prog)
    eng = D.Engine(D.SchedulerConfig(N.MODEL_RAFT5, model_flags=1))
    eng.set_externals(ext)
    res = eng.fuzz_batch(1, 2_000_000, 700, 100)
    seed = 1 + int(np.nonzero((res["violation"] == 1) & (res["steps"] >= 600))[0][0])
    ev, par, r = eng.fuzz_trace(seed, 700, 100)
    eng.set_trace(ev, ext)
    rng = np.random.default_rng(0)
    mw = eng.mask_words()
    t0 = time.perf_counter()
    bad = 0
    for c in range(0, n_masks, 100_000):
        m = min(100_000, n_masks - c)
        dens = rng.random((m, 1))
        bits = rng.random((m, mw * 64)) < dens
        masks = np.packbits(bits, axis=1, bitorder="little").view(np.uint64) & O.full_mask(ext)[None, :]
        g = eng.replay_batch(masks, int(r["violation"]))
        cpu = O.replay_batch(N.MODEL_RAFT5, ev, ext, masks, looking_for=int(r["violation"]), model_flags=1)
        bad += int((g != cpu).sum())
    return {"workload": "STSSched replays of random subsequences, %d-event trace, %d externals" % (len(ev), len(ext)),
            "replays": n_masks, "mismatching_records": bad, "seconds": round(time.perf_counter() - t0, 1)}


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    rows = [
        sweep("raft5 depth-50 interval-5 (bench workload, lane engine)", N.MODEL_RAFT5, D.raft5_program(), 1, 50, 5, n, 1_000_000),
        sweep("raft5 depth-50, pending multiset in the state hash", N.MODEL_RAFT5, D.raft5_program(client_cmds=3), 3, 50, 5,
              n // 10, 500_000, flags=1),
        sweep("raft5 depth-50 SrcDstFIFO", N.MODEL_RAFT5, D.raft5_program(), 1, 50, 5, n // 20, 250_000, strategy=1),
        sweep("pingpong3, 100 pings, to quiescence", N.MODEL_PINGPONG3, D.pingpong3_program(100), 0, -1, 0, n // 20, 250_000),
        sweep("bcast32 ttl-3 depth-200", N.MODEL_BCAST32, D.bcast32_program(3), 0, 200, 0, n // 50, 100_000, flags=1),
    ]
    cores = len(os.sched_getaffinity(0))
    rows.append(sweep_replay(n // 10))
    rows.append(sweep_dpor(2048, 200, min(cores, 32)))
    for r in rows:
        print(json.dumps(r))
    assert all(r.get("mismatching_records", 0) == 0 and r.get("mismatching_searches", 0) == 0 for r in rows)

