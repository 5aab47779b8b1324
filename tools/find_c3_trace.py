"""configs[3] needs a violating raft5 execution whose EventTrace has >= 2000 events: find one (deterministic, so the
seed it prints is then fixed in tools/secondary.py)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import demi_b200 as D
from demi_b200 import _native as N
from tools import secondary as S

eng = D.Engine(D.SchedulerConfig(N.MODEL_RAFT5, model_flags=1))
ext = D.pack_externals(D.raft5_program(client_cmds=S.C3_PROGRAM_CMDS))
eng.set_externals(ext)
n = 400_000
res = eng.fuzz_batch(1, n, S.C3_MAX_MESSAGES, S.C3_INTERVAL)
hits = np.nonzero((res["violation"] == 1) & (res["steps"] >= 900))[0]
print("violating prefixes with >= 900 deliveries:", len(hits), "of", n, "status!=0:", int((res["status"] != 0).sum()))
for i in hits[:20]:
    ev, par, r = eng.fuzz_trace(1 + int(i), S.C3_MAX_MESSAGES, S.C3_INTERVAL)
    print("seed", 1 + int(i), "events", len(ev), "steps", int(r["steps"]), "violation", int(r["violation"]))
    if len(ev) >= 2000:
        print("C3_SEED =", 1 + int(i))
        break
