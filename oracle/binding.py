"""ctypes binding of the CPU oracle (oracle/liboracle.so).

ORACLE — test infrastructure only.  Import this from tests/, from
__graft_entry__.smoke() and from bench.py's cpu_baseline / --impl reference
legs only; never from the demi_b200 package.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")

EXT_DTYPE = np.dtype([("kind", "u1"), ("a", "u1"), ("b", "u1"), ("type", "u1"),
                      ("p0", "<u4"), ("p1", "<u4"), ("id", "<u4")])
EVENT_DTYPE = np.dtype([("kind", "u1"), ("src", "u1"), ("dst", "u1"), ("type", "u1"),
                        ("p0", "<u4"), ("p1", "<u4"), ("uniq", "<u2"), ("node", "<u2")])
RESULT_DTYPE = np.dtype([("violation", "<u4"), ("steps", "<u4"), ("state_hash", "<u8"), ("trace_hash", "<u8"),
                         ("n_nodes", "<u2"), ("n_events", "<u2"), ("max_pending", "<u2"), ("status", "<u2")])


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("model", C.c_int32), ("model_flags", C.c_uint32),
                ("blocked_mask", C.c_uint32), ("ignore_timers", C.c_int32), ("reserved", C.c_int32 * 3)]


class FuzzParams(C.Structure):
    _fields_ = [("seed_base", C.c_int64), ("n_prefixes", C.c_uint64), ("max_messages", C.c_int32),
                ("invariant_check_interval", C.c_int32), ("looking_for", C.c_uint32), ("reserved", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("oracle library missing: run `make -C oracle`")
        _lib = C.CDLL(LIB_PATH)
    return _lib


def fuzz_batch(model, ext, seed_base, n, max_messages, interval, model_flags=0, blocked_mask=0,
               ignore_timers=0, looking_for=0, threads=None):
    ext = np.ascontiguousarray(ext, dtype=EXT_DTYPE)
    cfg = Config(0, model, model_flags, blocked_mask, ignore_timers)
    p = FuzzParams(seed_base, n, max_messages, interval, looking_for, 0)
    out = np.empty(n, dtype=RESULT_DTYPE)
    threads = threads or (os.cpu_count() or 1)
    rc = lib().oracle_fuzz_batch(C.byref(cfg), C.c_void_p(ext.ctypes.data), C.c_uint32(len(ext)), C.byref(p),
                                 C.c_void_p(out.ctypes.data), C.c_int(threads))
    if rc != 0:
        raise RuntimeError("oracle_fuzz_batch failed: %d" % rc)
    return out


def fuzz_trace(model, ext, seed, max_messages, interval, model_flags=0, blocked_mask=0, ignore_timers=0,
               looking_for=0, cap_events=65536, cap_nodes=65536):
    ext = np.ascontiguousarray(ext, dtype=EXT_DTYPE)
    cfg = Config(0, model, model_flags, blocked_mask, ignore_timers)
    p = FuzzParams(seed, 1, max_messages, interval, looking_for, 0)
    ev = np.zeros(cap_events, dtype=EVENT_DTYPE)
    par = np.zeros(cap_nodes, dtype=np.uint16)
    ne, nn = C.c_uint32(), C.c_uint32()
    res = np.zeros(1, dtype=RESULT_DTYPE)
    lib().oracle_fuzz_trace(C.byref(cfg), C.c_void_p(ext.ctypes.data), C.c_uint32(len(ext)), C.byref(p),
                            C.c_int64(seed), C.c_void_p(ev.ctypes.data), C.c_uint32(cap_events), C.byref(ne),
                            C.c_void_p(par.ctypes.data), C.c_uint32(cap_nodes), C.byref(nn),
                            C.c_void_p(res.ctypes.data))
    return ev[:ne.value].copy(), par[:nn.value].copy(), res[0]


def kat_jrandom(seed, bound, n):
    out = (C.c_int32 * n)()
    lib().oracle_kat_jrandom(C.c_int64(seed), C.c_int32(bound), C.c_int32(n), out)
    return list(out)


def kat_hashset(seed, ops, blocked_mask=0):
    ops_a = (C.c_int32 * len(ops))(*ops)
    arr = (C.c_int32 * (len(ops) + 1))()
    removed = (C.c_int32 * (len(ops) + 1))()
    nr = C.c_int32()
    n = lib().oracle_kat_hashset(C.c_int64(seed), C.c_uint32(blocked_mask), ops_a, C.c_int32(len(ops)),
                                 arr, removed, C.byref(nr))
    return list(arr[:n]), list(removed[:nr.value])
