"""world_size-2 gloo test of the N>1 path (host logic only; the per-rank compute is the oracle here)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_per_rank, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from demi_b200 import sharding, events
    from oracle import binding as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ext = events.pack_externals(events.raft5_program())
    base = sharding.seed_range(0, rank, world, n_per_rank)
    res = O.fuzz_batch(2, ext, base, n_per_rank, 50, 5, model_flags=1, threads=2)
    hits = np.nonzero(res["violation"])[0]
    seeds, codes = sharding.gather_violations(base + hits, res["violation"][hits])
    t = sharding.max_over_ranks(1.0 + rank)
    first, count = sharding.split_units(1001, rank, world)
    q.put((rank, seeds.tolist(), codes.tolist(), t, first, count))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank(oracle):
    from demi_b200 import events, sharding
    world, n_per_rank = 2, 3000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_per_rank, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # both ranks see the same union, equal to a single-rank run over the whole seed range
    ext = events.pack_externals(events.raft5_program())
    one = oracle.fuzz_batch(2, ext, 1, world * n_per_rank, 50, 5, model_flags=1)
    hits = np.nonzero(one["violation"])[0]
    assert outs[0][1] == outs[1][1] == (1 + hits).tolist()
    assert outs[0][2] == outs[1][2] == one["violation"][hits].tolist()
    assert outs[0][3] == outs[1][3] == 2.0                     # max over ranks
    assert (outs[0][4], outs[0][5], outs[1][4], outs[1][5]) == (0, 501, 501, 500)
    assert sharding.seed_range(2, 1, 4, 10) == 1 + (2 * 4 + 1) * 10
