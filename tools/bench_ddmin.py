"""configs[3] DDMin wall-clock: demi_ddmin on the 2039-event trace vs the sequential CPU oracle."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import demi_b200 as D
from demi_b200 import _native as N
from tools import secondary as S
from oracle import binding as O
eng, ext, ev, par, r, seed = S.c3_trace(0)
code = int(r["violation"])
eng.set_trace(ev, ext)
import os
for wide in (1500, 6000, 24000, 6000):
  os.environ['DEMI_DDMIN_WIDE'] = str(wide)
  for rep in range(2):
    t0 = time.perf_counter(); mcs, iters, dd = eng.ddmin(code); dt = time.perf_counter() - t0
    print(wide, "gpu ddmin %.2f ms: mcs %d, sequential tests %d, executed %d in %d batches" % (dt * 1e3, dd.mcs_size, dd.total_replays, dd.replays_executed, dd.batches), "host build %d us, evaluate %d us" % (dd.reserved[0], dd.reserved[1]), "last kernel %.2f ms" % eng.stats().kernel_ms)
t0 = time.perf_counter(); rc, cmcs, total, citers, ver = O.ddmin_sts(N.MODEL_RAFT5, ev, ext, code, model_flags=1); cdt = time.perf_counter() - t0
print("cpu oracle %.2f ms, identical %s" % (cdt * 1e3, bool((mcs == cmcs).all() and dd.total_replays == total and list(iters) == list(citers))))
