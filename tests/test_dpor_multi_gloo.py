"""world_size-2 gloo test of the DPOR work-stealing path (host logic; per-rank compute = the oracle)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _programs():
    import demi_b200 as D
    rng = np.random.default_rng(5)
    progs = []
    for i in range(24):
        ev = [D.Start(int(a)) for a in rng.permutation(5)]
        # the first half of the list is deliberately heavier than the second (more boots => longer searches)
        k = 5 if i < 12 else 2
        ev += [D.Send(int(a), 1, 0x1F) for a in rng.permutation(5)[:k]]
        progs.append(D.pack_externals(ev))
    return progs


def _oracle_batch(exts):
    from demi_b200 import _native as N
    from oracle import binding as O
    out = np.zeros(len(exts), dtype=N.DPOR_RESULT_DTYPE)
    for i, e in enumerate(exts):
        rc, r, _, _ = O.dpor_search(2, e, 16, 60, model_flags=1, node_cap=2048, explored_slots=1 << 14, heap_cap=1 << 14)
        out[i] = r
    return out


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from demi_b200 import dpor_multi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res, stats = dpor_multi.run_searches(_programs(), _oracle_batch, chunk=3)
    q.put((rank, res.tobytes(), stats))
    dist.barrier()
    dist.destroy_process_group()


def test_transfer_plan_levels_queues():
    sys.path.insert(0, ROOT)
    from demi_b200.dpor_multi import transfer_plan
    assert transfer_plan([10, 0]) == [(0, 1, 5)]
    assert transfer_plan([3, 3, 3]) == []
    plan = transfer_plan([9, 1, 0, 2])
    c = [9, 1, 0, 2]
    for s, d, k in plan:
        c[s] -= k
        c[d] += k
    assert sorted(c) == [3, 3, 3, 3]
    assert transfer_plan([0, 0]) == []


def test_two_rank_work_stealing_matches_single_rank(oracle):
    from demi_b200 import _native as N
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = _oracle_batch(_programs())
    for rank, blob, stats in outs:
        got = np.frombuffer(blob, dtype=N.DPOR_RESULT_DTYPE)
        assert (got == single).all()
    assert outs[0][2]["executed"] + outs[1][2]["executed"] == 24
    assert outs[0][2]["sent_out"] + outs[1][2]["sent_out"] == outs[0][2]["stolen_in"] + outs[1][2]["stolen_in"]
