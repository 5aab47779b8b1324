"""Multi-GPU sharding of the schedule-exploration workloads.

The reference runs executions strictly one after another in one JVM
(RandomScheduler.scala:248-269; one ActorSystem at a time, Instrumenter.scala:
203-218); fuzz prefixes and DDMin tests are independent units, so ranks take
disjoint unit ranges and never exchange data on the data path.  The only
collectives are the final gather of (small) violation records and the
max-over-ranks reduction of the timings.
"""
import numpy as np
import torch
import torch.distributed as dist


def seed_range(step, rank, world, n_per_rank, seed0=1):
    """First seed of the `n_per_rank` consecutive seeds rank `rank` explores at step `step`."""
    return seed0 + (step * world + rank) * n_per_rank


def split_units(n_total, rank, world):
    """Block partition of n_total units: (first, count) for `rank`."""
    base, rem = divmod(n_total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def gather_violations(seeds, codes, device=None, max_per_rank=1 << 20):
    """All ranks contribute (seed, violation code) pairs; every rank receives the
    sorted union (the violating-schedule set of the whole job)."""
    seeds = np.asarray(seeds, dtype=np.int64)
    codes = np.asarray(codes, dtype=np.int64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        order = np.argsort(seeds, kind="stable")
        return seeds[order], codes[order]
    world = dist.get_world_size()
    dev = device if device is not None else torch.device("cpu")
    n = torch.tensor([min(len(seeds), max_per_rank)], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, n)
    cap = int(max(int(c.item()) for c in counts))
    buf = torch.zeros((max(cap, 1), 2), dtype=torch.int64, device=dev)
    k = int(n.item())
    if k:
        buf[:k, 0] = torch.from_numpy(seeds[:k]).to(dev)
        buf[:k, 1] = torch.from_numpy(codes[:k]).to(dev)
    outs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    parts = [o[:int(c.item())].cpu().numpy() for o, c in zip(outs, counts)]
    allp = np.concatenate(parts, axis=0) if parts else np.zeros((0, 2), dtype=np.int64)
    order = np.argsort(allp[:, 0], kind="stable")
    return allp[order, 0], allp[order, 1]


def max_over_ranks(value, device=None):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else torch.device("cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
