"""BASELINE configs[2] on N GPUs: independent DPOR searches with NCCL work-stealing rebalance.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_dpor_multi.py"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import demi_b200 as D
from demi_b200 import _native as N, dpor_multi

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

rng = np.random.default_rng(11)
progs = []
n_total = 16384 * world          # 8192-search launches
for i in range(n_total):
    ev = [D.Start(int(a)) for a in rng.permutation(5)]
    # skewed: early programs boot all five nodes (long searches), late ones a single node (short searches)
    k = 5 if i < n_total // 2 else 1
    ev += [D.Send(int(a), 1, 0x1F) for a in rng.permutation(5)[:k]]
    progs.append(D.pack_externals(ev))
eng = D.Engine(D.SchedulerConfig(N.MODEL_RAFT5, model_flags=3, device=lr))


def run_batch(exts):
    res, _, _ = eng.dpor_batch(exts, 100, 60, node_cap=2048, explored_slots=1 << 16, heap_cap=1 << 16)
    return res


run_batch(progs[:32])
out = {}
skew = [3] + [1] * (world - 1)          # three quarters of the searches start on rank 0
for rebalance in (False, True):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res, stats = dpor_multi.run_searches(progs, run_batch, chunk=8192, device=dev, rebalance=rebalance, initial_weights=skew)
    torch.cuda.synchronize()
    dt = dpor_multi_max = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    out["rebalance" if rebalance else "static"] = {
        "seconds": dt, "interleavings": int(res["interleavings"].sum()), "interleavings_per_s": int(res["interleavings"].sum()) / dt,
        "rank0_executed": stats["executed"], "rank0_stolen_in": stats["stolen_in"], "rank0_sent_out": stats["sent_out"],
        "rounds": stats["rounds"], "status_ok": bool((res["status"] == 0).all()), "checksum": int(res["n_explored"].astype(np.uint64).sum())}
if rank == 0:
    print(json.dumps({"config": "configs[2]: raft5 DPORwHeuristics depth-100, %d searches, %d GPU(s)" % (n_total, world), **out}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
