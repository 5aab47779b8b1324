"""The C-ABI library loads and exports every symbol include/demi_b200.h declares
(no compute calls: there is no GPU on the CPU runner)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "demi_b200.h")).read()
    return sorted(set(re.findall(r"\b(demi_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree(native):
    assert declared_symbols() == sorted(native.EXPORTS)


def test_library_exports_every_declared_symbol(native):
    L = native.lib()
    for name in declared_symbols():
        assert hasattr(L, name), name
    assert b"sm_100a" in L.demi_version()


def test_struct_sizes(native):
    assert C.sizeof(native.Config) == 32
    assert C.sizeof(native.FuzzParams) == 32
    assert native.RESULT_DTYPE.itemsize == 32 and native.EVENT_DTYPE.itemsize == 16 and native.EXT_DTYPE.itemsize == 16


def test_no_cpu_fallback(native):
    """Without a CUDA device the product path fails loudly instead of computing on the host."""
    import demi_b200 as D
    if native.lib().demi_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(D.DemiError) as ei:
        D.Engine(D.SchedulerConfig(native.MODEL_RAFT5))
    assert ei.value.code == native.ERR_NO_DEVICE


def test_create_rejects_bad_arguments(native):
    L = native.lib()
    h = C.c_void_p()
    assert L.demi_create(None, C.byref(h)) == native.ERR_INVALID
    cfg = native.Config(0, 99, 0, 0, 0)
    assert L.demi_create(C.byref(cfg), C.byref(h)) == native.ERR_INVALID
    assert b"unknown model" in L.demi_last_error(None)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under demi_b200/ may reference it."""
    for d, _, files in os.walk(os.path.join(ROOT, "demi_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(d, f)).read()
                if f == "build.py":
                    continue        # builds the checker library; does not load it
                # (the word itself is the reference's vocabulary: TestOracle, DDMin(oracle, ...))
                for needle in ("import oracle", "from oracle", "liboracle", "oracle/", "oracle.binding", "oracle_"):
                    assert needle not in txt, (os.path.join(d, f), needle)


def test_external_event_packing_roundtrip():
    import demi_b200 as D
    prog = D.raft5_program(client_cmds=2) + [D.Partition(0, 1), D.UnPartition(0, 1), D.Kill(3)]
    arr = D.pack_externals(prog)
    back = D.unpack_externals(arr)
    assert [type(a) for a in back] == [type(a) for a in prog]
    assert back == prog                                # UniqueExternalEvent equality is by id
    assert (D.pack_externals(back) == arr).all()
    assert len({e._id for e in prog}) == len(prog)
