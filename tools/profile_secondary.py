"""Small drivers for ncu captures of K2 (replay), K3 (DPOR), K4/K5 (dedup/compact)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import demi_b200 as D
from demi_b200 import _native as N
from oracle import binding as O

which = sys.argv[1]
if which == "replay":
    prog = D.raft5_program(client_cmds=290)
    ext = D.pack_externals(prog)
    eng = D.Engine(D.SchedulerConfig(N.MODEL_RAFT5, model_flags=1))
    eng.set_externals(ext)
    res = eng.fuzz_batch(1, 2_000_000, 700, 100)
    seed = 1 + int(np.nonzero((res["violation"] == 1) & (res["steps"] >= 600))[0][0])
    ev, par, r = eng.fuzz_trace(seed, 700, 100)
    eng.set_trace(ev, ext)
    rng = np.random.default_rng(0)
    n, mw = 1_000_000, eng.mask_words()
    masks = rng.integers(0, 2**63, size=(n, mw), dtype=np.uint64) * np.uint64(2) | (rng.integers(0, 2**63, size=(n, mw), dtype=np.uint64) * np.uint64(2))
    masks &= O.full_mask(ext)[None, :]
    for _ in range(2):
        eng.replay_batch(masks, int(r["violation"]))
elif which == "dpor":
    rng = np.random.default_rng(7)
    progs = []
    for _ in range(4096):
        ev = [D.Start(int(a)) for a in rng.permutation(5)]
        ev += [D.Send(int(a), 1, 0x1F) for a in rng.permutation(5)[:int(rng.integers(3, 6))]]
        progs.append(ev)
    eng = D.Engine(D.SchedulerConfig(N.MODEL_RAFT5, model_flags=3))
    eng.dpor_batch(progs, 100, 40, heap_cap=1 << 16)
elif which == "dedup":
    eng = D.Engine(D.SchedulerConfig(N.MODEL_RAFT5))
    m = 20_000_000
    rec = np.zeros(m, dtype=N.RESULT_DTYPE)
    rec["state_hash"] = np.random.default_rng(1).integers(0, 1_000_000, size=m, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    big = torch.from_numpy(rec.view(np.uint8)).cuda()
    bout = torch.empty_like(big)
    bidx = torch.empty(m, dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    for _ in range(2):
        eng.dedup_compact_dev(big.data_ptr(), m, 0, bout.data_ptr(), bidx.data_ptr(), cnt.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
elif which == "prov":
    eng = D.Engine(D.SchedulerConfig(N.MODEL_RAFT5, model_flags=1))
    eng.set_externals(D.raft5_program())
    res = eng.fuzz_batch(1, 1_000_000, 50, 5)
    viol = np.nonzero(res["violation"])[0].astype(np.uint32)
    eng.fuzz_provenance(1, viol, 50, 5)
elif which == "dpor_arvind":
    eng = D.Engine(D.SchedulerConfig(N.MODEL_RAFT5, model_flags=1))
    prog = D.raft5_program()
    eng.set_externals(prog)
    ext_all = D.pack_externals(prog)
    dext = ext_all[(ext_all["kind"] == 1) | (ext_all["kind"] == 3)]
    res = eng.fuzz_batch(1, 3000, 40, 5)
    i = int(np.nonzero(res["violation"] == 1)[0][3])
    ev, par, r = eng.fuzz_trace(1 + i, 40, 5)
    rng = np.random.default_rng(3)
    progs, caps = [], []
    for t in range(4096):
        keep = np.ones(len(dext), dtype=bool)
        keep[rng.integers(0, len(dext), size=int(rng.integers(0, 4)))] = False
        progs.append(dext[keep]); caps.append([0, 2, 4, 8, 16, 32][:int(rng.integers(1, 7))])
    eng.dpor_batch_ex(progs, int(r["steps"]), 64, seed=(ev, par), flags=3, caps=caps, looking_for=1, heap_cap=1 << 14)
print("done", which)
