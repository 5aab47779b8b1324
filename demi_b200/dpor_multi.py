"""Multi-GPU DPOR: independent searches, rank-local queues, periodic work-stealing rebalance.

One DPORwHeuristics search is sequential by definition (see dpor_kernel.cuh); the
shardable unit is the search (one per external-event program).  Searches differ
wildly in length (a search may exhaust its backtrack set after a few dozen
interleavings or run into the budget), so a static partition leaves ranks idle.
Each rank therefore works through its queue in rounds; after every round the
ranks all-gather their queue lengths (8 bytes each), every rank computes the same
deterministic transfer plan, and surplus *pending searches* (their external
programs: a few hundred bytes each) move from loaded to idle ranks in one
all-to-all over NVLink.  Results do not depend on where a search ran, so the
merged output is identical to a single-rank run.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _native as N


def transfer_plan(counts):
    """Deterministic plan [(src, dst, k), ...] that levels `counts` to within one unit."""
    world = len(counts)
    total = sum(counts)
    base, rem = divmod(total, world)
    target = [base + (1 if r < rem else 0) for r in range(world)]
    surplus = [(r, counts[r] - target[r]) for r in range(world) if counts[r] > target[r]]
    deficit = [(r, target[r] - counts[r]) for r in range(world) if counts[r] < target[r]]
    plan, i, j = [], 0, 0
    while i < len(surplus) and j < len(deficit):
        s, sk = surplus[i]
        d, dk = deficit[j]
        k = min(sk, dk)
        plan.append((s, d, k))
        sk -= k
        dk -= k
        surplus[i] = (s, sk)
        deficit[j] = (d, dk)
        if sk == 0:
            i += 1
        if dk == 0:
            j += 1
    return plan


def _pack(items):
    """[(gid, ext records)] -> uint8 buffer: int64 n, then per item int64 gid, int64 len, records."""
    parts = [np.array([len(items)], dtype=np.int64).view(np.uint8)]
    for gid, ext in items:
        parts.append(np.array([gid, len(ext)], dtype=np.int64).view(np.uint8))
        parts.append(np.ascontiguousarray(ext, dtype=N.EXT_DTYPE).view(np.uint8).reshape(-1))
    return np.concatenate(parts)


def _unpack(buf):
    if len(buf) == 0:
        return []
    n = int(buf[:8].view(np.int64)[0])
    off, out = 8, []
    for _ in range(n):
        gid, ln = (int(x) for x in buf[off:off + 16].view(np.int64))
        off += 16
        ext = buf[off:off + ln * 16].view(N.EXT_DTYPE).copy()
        off += ln * 16
        out.append((gid, ext))
    return out


def _exchange(outgoing, device):
    """outgoing[r] = uint8 array for rank r; returns incoming[r] from every rank."""
    world = dist.get_world_size()
    sizes = torch.tensor([len(o) for o in outgoing], dtype=torch.int64, device=device)
    in_sizes = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(in_sizes, sizes)
    in_list = [int(x) for x in in_sizes.cpu()]
    send = torch.from_numpy(np.concatenate(outgoing) if sum(len(o) for o in outgoing) else np.zeros(0, np.uint8)).to(device)
    recv = torch.empty(sum(in_list), dtype=torch.uint8, device=device)
    dist.all_to_all_single(recv, send, output_split_sizes=in_list, input_split_sizes=[len(o) for o in outgoing])
    recv = recv.cpu().numpy()
    out, off = [], 0
    for k in in_list:
        out.append(recv[off:off + k])
        off += k
    return out


def run_searches(programs, run_batch, chunk=256, device=None, rebalance=True, quantum_s=0.0, initial_weights=None):
    """programs: the FULL list of external programs (same on every rank).
    run_batch(list of ext arrays) -> structured array of demi_dpor_result, one per program.
    A round = chunks until `quantum_s` seconds have passed (at least one chunk); ranks whose
    searches are short drain their queues faster, which is what the rebalance then evens out.
    Returns (results for all programs in program order [identical on every rank], stats dict)."""
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    world = dist.get_world_size() if distributed else 1
    rank = dist.get_rank() if distributed else 0
    dev = device if device is not None else torch.device("cpu")
    packed = [np.ascontiguousarray(p, dtype=N.EXT_DTYPE) for p in programs]
    # initial partition: contiguous blocks, optionally skewed (`initial_weights`, e.g. searches arriving
    # at one rank); stealing has to fix the imbalance
    w = list(initial_weights) if initial_weights else [1] * world
    cum = np.concatenate([[0], np.cumsum(w)]) / float(sum(w))
    lo = int(round(len(packed) * cum[rank]))
    hi = int(round(len(packed) * cum[rank + 1]))
    queue = [(g, packed[g]) for g in range(lo, hi)]
    done = {}
    stats = {"rounds": 0, "stolen_in": 0, "sent_out": 0, "executed": 0}
    import time
    while True:
        t_round = time.perf_counter()
        while queue:
            batch, queue = queue[:chunk], queue[chunk:]
            res = run_batch([e for _, e in batch])
            for (g, _), r in zip(batch, res):
                done[g] = r
            stats["executed"] += len(batch)
            if time.perf_counter() - t_round >= quantum_s:
                break
        stats["rounds"] += 1
        if not distributed:
            if not queue:
                break
            continue
        counts_t = torch.tensor([len(queue)], dtype=torch.int64, device=dev)
        gathered = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(gathered, counts_t)
        counts = [int(x.item()) for x in gathered]
        if sum(counts) == 0:
            break
        if not rebalance:
            continue
        plan = transfer_plan(counts)
        if not plan:
            continue
        outgoing = [[] for _ in range(world)]
        for s, d, k in plan:
            if s == rank:
                give, queue = queue[len(queue) - k:], queue[:len(queue) - k]   # surplus leaves from the tail
                outgoing[d].extend(give)
                stats["sent_out"] += k
        incoming = _exchange([_pack(o) if o else np.zeros(0, np.uint8) for o in outgoing], dev)
        for buf in incoming:
            got = _unpack(buf)
            queue.extend(got)
            stats["stolen_in"] += len(got)
    # merge: every rank ends with every result, in program order
    n = len(packed)
    mine = np.zeros(n, dtype=N.DPOR_RESULT_DTYPE)
    have = np.zeros(n, dtype=np.uint8)
    for g, r in done.items():
        mine[g] = r
        have[g] = 1
    if distributed:
        t = torch.from_numpy(mine.view(np.uint8).reshape(n, -1).copy()).to(dev)
        hmask = torch.from_numpy(have.copy()).to(dev)
        ts = [torch.zeros_like(t) for _ in range(world)]
        hs = [torch.zeros_like(hmask) for _ in range(world)]
        dist.all_gather(ts, t)
        dist.all_gather(hs, hmask)
        out = np.zeros(n, dtype=N.DPOR_RESULT_DTYPE)
        seen = np.zeros(n, dtype=np.int64)
        for tt, hh in zip(ts, hs):
            a = tt.cpu().numpy().reshape(-1).view(N.DPOR_RESULT_DTYPE)
            m = hh.cpu().numpy().astype(bool)
            out[m] = a[m]
            seen += m
        assert (seen == 1).all(), "every search must run exactly once"
        return out, stats
    return mine, stats
