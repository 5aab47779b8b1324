"""A plain-C host (gcc + dlopen, no Python, no C++ mirror) drives fuzz -> trace -> DDMin -> verify -> frontier DPOR
through include/demi_b200.h: what the JNI shim forwards to."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "abi_pipeline.c")
EXE = os.path.join(ROOT, "tests", "c", "abi_pipeline")
LIB = os.path.join(ROOT, "demi_b200", "libdemi_b200.so")


def build(native):
    native.lib()
    subprocess.check_call(["gcc", "-std=gnu11", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), SRC, "-o", EXE, "-ldl"])


def test_c_host_builds_and_sees_the_no_fallback_contract(native):
    build(native)
    if native.lib().demi_device_count() > 0:
        pytest.skip("a CUDA device is present (covered by the gpu test)")
    out = subprocess.run([EXE, LIB], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "no-fallback contract" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_c_host_pipeline_on_gpu(native):
    build(native)
    out = subprocess.run([EXE, LIB], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr
