"""configs[2]: ONE raft5 DPORwHeuristics search at depth 100, frontier engine (demi_dpor_frontier), 1 GPU."""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np
import demi_b200 as D
from demi_b200 import _native as N

def main():
    width = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    budget = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
    eng = D.Engine(D.SchedulerConfig(N.MODEL_RAFT5, model_flags=3))
    prog = D.raft5_program(client_cmds=2)[:-1]
    flags = N.FR_NO_HISTORY if (len(sys.argv) > 3 and sys.argv[3] == "nohist") else 0
    F = eng.frontier_params(100, budget, width, explored_slots=1 << 24, pool_cap=(1 << 30) if flags else (1 << 24),
                            trace_cap=budget + width + 16, flags=flags)
    for rep in range(2):
        t = time.perf_counter()
        r, viol, hashes = eng.dpor_frontier(prog, F, cap_viol=1 << 20, want_hashes=False)
        dt = time.perf_counter() - t
        out = {k: (float(r[k]) if r[k].dtype.kind == "f" else int(r[k])) for k in r.dtype.names}
        out.update(width=width, wall_s=dt, interleavings_per_s=int(r["interleavings"]) / dt,
                   kernel_interleavings_per_s=int(r["interleavings"]) / ((r["exec_ms"] + r["scan_ms"] + r["select_ms"]) * 1e-3))
        print(json.dumps(out))

if __name__ == "__main__":
    main()
