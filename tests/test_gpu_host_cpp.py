"""The C++ host mirror (zig-tfhe_b200/host/tfhe_b200.hpp) end to end: compiled with g++ against the C-ABI
library, run on the GPU, compared with the oracle."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ROOT, keys_for
from oracle import oracle as O


def _build(tmp):
    exe = os.path.join(tmp, "host_cpp_smoke")
    lib_dir = os.path.join(ROOT, "zig-tfhe_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "host_cpp_smoke.cpp"),
                           "-L" + lib_dir, "-ltfhe_b200", "-Wl,-rpath," + lib_dir])
    return exe


def test_cpp_host_mirror_compiles_and_fails_loudly_without_gpu():
    import torch
    with tempfile.TemporaryDirectory() as tmp:
        exe = _build(tmp)
        if torch.cuda.is_available():
            pytest.skip("GPU present: covered by the gpu test")
        for name, arr in (("bsk", np.zeros(8)), ("ksk", np.zeros(8, np.uint32)), ("a", np.zeros(701, np.uint32)), ("b", np.zeros(701, np.uint32))):
            arr.tofile(os.path.join(tmp, name + ".bin"))
        r = subprocess.run([exe, tmp], capture_output=True, text=True)
        assert r.returncode == 4 and "error 2" in r.stdout      # TFHE_B200_ERR_NO_DEVICE, no CPU fallback


@pytest.mark.gpu
def test_cpp_host_mirror_on_gpu():
    orc = O.Oracle("128"); keys = keys_for("128")
    bits = np.array([1, 0, 1], np.uint8)
    ca = orc.encrypt_bools(bits, keys, 61); cb = orc.encrypt_bools(np.array([1, 1, 0], np.uint8), keys, 62)
    with tempfile.TemporaryDirectory() as tmp:
        exe = _build(tmp)
        keys.bsk.tofile(os.path.join(tmp, "bsk.bin")); keys.ksk.tofile(os.path.join(tmp, "ksk.bin"))
        ca.tofile(os.path.join(tmp, "a.bin")); cb.tofile(os.path.join(tmp, "b.bin"))
        r = subprocess.run([exe, tmp], capture_output=True, text=True)
        assert r.returncode == 0 and "strategy=b200" in r.stdout, r.stdout + r.stderr
        out = np.fromfile(os.path.join(tmp, "out.bin"), np.uint32).reshape(-1, 701)
    assert (out[:3] == orc.gate_batch(O.NAND, ca, cb, keys)).all()
    mux = orc.gate(O.OR, orc.gate(O.AND, ca[0], cb[0], keys), orc.gate(O.AND, orc.gate_not(ca[0]), ca[1], keys), keys)
    assert (out[3] == mux).all()
    assert (out[4] == orc.bootstrap(ca[0], keys)).all()
    # Circuit::rippleCarryAdder(1) over the 3 instances with cin = NOT a: fullAdder gate by gate (add_two_numbers.zig:24-39)
    cin = np.stack([orc.gate_not(ca[i]) for i in range(3)])
    axb = orc.gate_batch(O.XOR, ca, cb, keys); ab = orc.gate_batch(O.AND, ca, cb, keys)
    t = orc.gate_batch(O.AND, axb, cin, keys)
    assert (out[5:8] == orc.gate_batch(O.XOR, axb, cin, keys)).all()
    assert (out[8:11] == orc.gate_batch(O.OR, ab, t, keys)).all()
