"""Regenerates the committed fixtures from the CPU oracle (there are no reference-side golden
vectors: DEMi has no tests, and it cannot run here — see DESIGN.md §5).  The fixtures pin the
oracle itself against accidental drift and give the GPU tests a second, frozen comparison point.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from demi_b200 import events as E          # noqa: E402
from oracle import binding as O            # noqa: E402

CASES = {
    # name: (model, program, model_flags, seed_base, n, max_messages, interval)
    "raft5_bug_d50": (2, E.raft5_program(), 1, 1, 2000, 50, 5),
    "raft5_clients_d100": (2, E.raft5_program(client_cmds=3), 3, 1000, 500, 100, 30),
    "pingpong3_c1": (1, E.pingpong3_program(100), 0, 1, 200, -1, 0),
    "bcast32_ttl2": (3, E.bcast32_program(2), 0, 1, 50, 200, 0),
}


def main():
    for name, (model, prog, flags, seed, n, maxm, interval) in CASES.items():
        # external-event ids are process-global counters: normalise them so the fixture is stable
        ext = E.pack_externals(prog)
        ext["id"] = np.arange(1, len(ext) + 1)
        res = O.fuzz_batch(model, ext, seed, n, maxm, interval, model_flags=flags, threads=4)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), ext=ext, results=res,
                            params=np.array([model, flags, seed, n, maxm, interval], dtype=np.int64))
        print(name, "violations", int((res["violation"] != 0).sum()), "of", n)
    # one full EventTrace + dep tree + STS/DDMin outcome
    model, prog, flags = 2, E.raft5_program(client_cmds=6), 1
    ext = E.pack_externals(prog)
    ext["id"] = np.arange(1, len(ext) + 1)
    res = O.fuzz_batch(model, ext, 1, 5000, 50, 5, model_flags=flags, threads=4)
    seed = 1 + int(np.nonzero(res["violation"])[0][0])
    ev, par, r = O.fuzz_trace(model, ext, seed, 50, 5, model_flags=flags)
    rc, mcs, total, iters, ver = O.ddmin_sts(model, ev, ext, int(r["violation"]), model_flags=flags)
    rng = np.random.default_rng(0)
    masks = (rng.integers(0, 2 ** len(ext), size=256, dtype=np.uint64) & O.full_mask(ext)[0]).reshape(-1, 1)
    rep = O.replay_batch(model, ev, ext, masks, looking_for=int(r["violation"]), model_flags=flags, threads=4)
    np.savez_compressed(os.path.join(HERE, "raft5_trace_ddmin.npz"), ext=ext, seed=np.int64(seed), events=ev, dep_parent=par,
                        result=np.array([r]), mcs=mcs, total_replays=np.int64(total), iteration_sizes=iters,
                        verified=np.int64(ver), masks=masks, replay=rep)
    print("trace", len(ev), "events; MCS", bin(int(mcs[0])), "tests", total)
    extras(model, flags, ext, seed, ev, par, int(r["violation"]), int(r["steps"]))


def extras(model, flags, ext, seed, ev, par, code, steps):
    """Round-1 additions on the same recorded execution: provenance pruning, plain and seeded / capped DPOR,
    IncrementalDDMin, internal minimization with both removal strategies."""
    out = {}
    keep, po = O.fuzz_provenance(model, ext, seed, 50, 5, 1, model_flags=flags)
    out["prov_keep"], out["prov_out"] = keep, np.array([po])
    dext = ext[(ext["kind"] == 1) | (ext["kind"] == 3)]
    rc, dr, dv, dh = O.dpor_search(model, dext, 40, 60, model_flags=flags, node_cap=4096, explored_slots=1 << 16, heap_cap=1 << 16)
    out["dpor_result"], out["dpor_hashes"] = np.array([dr]), dh
    sd = O.dpor_seed(ev, par)
    rows, hashes = [], []
    for arvind, prio, caps in ((1, 1, [0, 2, 4, 8, -1]), (0, 1, [0, 0, -1]), (0, 0, [-1])):
        inst = O.DporInstance(model, dext, steps, 80, seed=sd, arvind=arvind, prioritize_pending=prio, model_flags=flags,
                              looking_for=code)
        for c in caps:
            rr, hh = inst.test(c)
            rows.append(rr); hashes.append(hh)
        inst.close()
    out["inst_results"] = np.array(rows)
    out["inst_hashes"] = np.concatenate(hashes) if hashes else np.zeros(0, dtype=np.uint64)
    rc, mcs, st = O.incremental_ddmin(model, dext, steps, 2000, sd, model_flags=flags, looking_for=code, stop_at_size=1,
                                      max_max_distance=64)
    out["inc_mcs"] = mcs
    out["inc_stats"] = np.array([st["total_replays"], st["rounds"], st["interleavings"], st["instances"]], dtype=np.int64)
    rr, vtrace = O.replay_trace(model, ev, ext, O.full_mask(ext), looking_for=code, model_flags=flags)
    mext = ext[ext["kind"] != 4]
    for name, fl in (("ltr", 0), ("fifo", 0x100)):
        rc, tr, total, sizes, unig = O.internal_minimize(model, vtrace, mext, code, model_flags=flags, flags=fl)
        out["im_%s_trace" % name], out["im_%s_sizes" % name] = tr, sizes
        out["im_%s_meta" % name] = np.array([rc, total, unig], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "raft5_round1_extras.npz"), **out)
    print("extras:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
