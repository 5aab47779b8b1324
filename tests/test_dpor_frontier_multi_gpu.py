"""K3F on several GPUs of one box: rank-local queues + the in-library NCCL steal round (demi_create_multi /
demi_dpor_frontier_multi, one host thread per device) against the CPU simulation of the same protocol."""
import ctypes as C

import numpy as np
import pytest

import demi_b200 as D
from demi_b200 import _native as N

pytestmark = pytest.mark.gpu

FIELDS = ["interleavings", "violations", "deliveries", "races", "keys_enqueued", "keys_dropped", "explored_pairs",
          "pool_left", "records_sent", "records_received", "bytes_sent", "rounds", "exchanges", "exhausted",
          "budget_exhausted", "status", "trace_slots"]


def run_multi(model, prog, flags, F, n, cap_viol=4096):
    L = N.lib()
    if L.demi_device_count() < n:
        pytest.skip("needs %d GPUs" % n)
    cfg = N.Config(0, model, flags, 0, 0, 0)
    devs = (C.c_int32 * n)(*range(n))
    hs = (C.c_void_p * n)()
    assert L.demi_create_multi(C.byref(cfg), devs, n, hs) == 0, L.demi_last_error(None)
    ext = D.pack_externals(prog)
    res = np.zeros(n, dtype=N.FRONTIER_RESULT_DTYPE)
    viol = np.zeros((n, cap_viol), dtype=N.DPOR_VIOL_DTYPE)
    cap_h = int(F.max_interleavings) + int(F.width) + 1
    hashes = np.zeros((n, cap_h), dtype=np.uint64)
    rc = L.demi_dpor_frontier_multi(hs, n, ext.ctypes.data, len(ext), C.byref(F), res.ctypes.data, viol.ctypes.data, cap_viol,
                                    hashes.ctypes.data, cap_h)
    err = [L.demi_last_error(hs[i]) for i in range(n)]
    for i in range(n):
        L.demi_destroy(hs[i])
    assert rc == 0, err
    return res, [viol[r, :int(res[r]["violations"])] for r in range(n)], [hashes[r, :int(res[r]["interleavings"])] for r in range(n)]


@pytest.mark.parametrize("n,width,S", [(2, 32, 2), (2, 256, 1), (4, 16, 2)])
def test_multi_rank_matches_the_simulated_protocol(oracle, n, width, S):
    prog = D.raft5_program(client_cmds=2)[:-1]
    for fr_flags, maxm, budget in [(0, 100, 100000), (N.FR_NO_HISTORY, 60, 6000)]:
        kw = dict(explored_slots=1 << 20, pool_cap=1 << 22, rounds_per_exchange=S, steal_max=128, flags=fr_flags)
        F = D.Engine.frontier_params(maxm, budget, width, **kw)
        res, viol, hashes = run_multi(N.MODEL_RAFT5, prog, 3, F, n)
        OF = oracle.frontier_params(maxm, budget, width, **kw)
        rc, ores, oviol, ohashes = oracle.dpor_frontier(N.MODEL_RAFT5, D.pack_externals(prog), OF, n, model_flags=3)
        assert rc == 0
        for r in range(n):
            for f in FIELDS:
                assert res[r][f] == ores[r][f], (r, f, res[r], ores[r])
            assert (hashes[r] == ohashes[r]).all()
            assert (viol[r] == oviol[r]).all()
        assert res["records_sent"].sum() == res["records_received"].sum() > 0
        assert (res["interleavings"] > 0).all()
