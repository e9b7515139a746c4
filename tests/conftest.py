"""Shared fixtures.  Tests marked `gpu` need a B200; everything else runs on CPU.

The oracle (oracle/) is the checker only: product code under zig-tfhe_b200/ never imports it."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zig-tfhe_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run on the B200 box)")


from oracle import oracle as O  # noqa: E402


def _ensure_native_artifacts():
    """A fresh checkout has no .so files (they are git-ignored): build them once (nvcc cross-compiles without a GPU)."""
    lib = os.path.join(ROOT, "zig-tfhe_b200", "libtfhe_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__ as entry
        entry.build_cuda()
    O.build()


_ensure_native_artifacts()


@pytest.fixture(scope="session")
def orc128():
    return O.Oracle("128")


_KEYS = {}


def keys_for(name, seed=1, with_ksk=True):
    k = (name, seed, with_ksk)
    if k not in _KEYS:
        _KEYS[k] = O.Oracle(name).keygen(seed=seed, with_ksk=with_ksk)
    return _KEYS[k]


@pytest.fixture(scope="session")
def keys128():
    return keys_for("128")


@pytest.fixture(scope="session")
def emu():
    d = os.path.join(ROOT, "tests", "emu")
    subprocess.check_call(["make", "-C", d], stdout=subprocess.DEVNULL)
    lib = C.CDLL(os.path.join(d, "libtfhe_emu.so"))
    return lib


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


TRUTH = {
    O.NAND: lambda a, b: 1 - (a & b),
    O.OR: lambda a, b: a | b,
    O.AND: lambda a, b: a & b,
    O.XOR: lambda a, b: a ^ b,
    # reference quirk: xnorGate computes a - 2b - 1/4 (gates.zig:78-82), which decrypts as XOR
    O.XNOR: lambda a, b: a ^ b,
    O.NOR: lambda a, b: 1 - (a | b),
    O.ANDNY: lambda a, b: (1 - a) & b,
    O.ANDYN: lambda a, b: a & (1 - b),
    O.ORNY: lambda a, b: (1 - a) | b,
    O.ORYN: lambda a, b: a | (1 - b),
}
