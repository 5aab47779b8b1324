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
    for r in rows:
        print(json.dumps(r))
    assert all(r["mismatching_records"] == 0 for r in rows)
