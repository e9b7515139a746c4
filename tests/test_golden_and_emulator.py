"""(1) The oracle against its committed golden fixture; (2) the host emulator of the CUDA kernel's
arithmetic (same __host__ __device__ code as the kernel) against the oracle, bit for bit."""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, keys_for, ptr
from oracle import oracle as O

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def test_oracle_matches_golden_fixture():
    import make_golden
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_128.json")))
    got = make_golden.build()
    assert got["sha256"] == want["sha256"]
    assert got["nand_out_row0"] == want["nand_out_row0"]
    assert got["nand_bits"] == want["nand_bits"] == [1, 1, 1, 0]
    assert got["offset"] == want["offset"] == 0x82080000
    assert got["ifft1024_head"] == want["ifft1024_head"]


def test_emulated_transform_matches_reference_spectrum(emu):
    rng = np.random.default_rng(3)
    poly = rng.integers(-32, 32, 1024).astype(np.int32)
    out = np.empty(1024, np.float64)
    emu.emu_forward(ptr(poly), ptr(out))
    Z = out.reshape(512, 2) @ np.array([1, 1j])
    ref = O.ifft1024(poly.view(np.uint32))
    R = (ref[:512] + 1j * ref[512:]) / 2.0
    t = np.arange(64); q2, q1 = t >> 3, t & 7
    for q0 in range(8):
        j = (512 - (q2 + 8 * q1 + 64 * q0)) & 511          # leaf (q2,q1,q0) = reference bin -(q2+8q1+64q0) mod 512
        assert np.abs(Z[q0 * 64 + t] - R[j]).max() < 1e-9 * max(1.0, np.abs(R).max())
    x = rng.integers(-2**31, 2**31, 1024).astype(np.int32)
    back = np.empty(1024, np.float64)
    emu.emu_forward(ptr(x), ptr(out)); emu.emu_inverse(ptr(out), ptr(back))
    assert np.abs(back / 512 - x).max() < 1e-4


@pytest.mark.parametrize("name,count", [("128", 2), ("80", 1), ("110", 1)])
def test_emulated_blind_rotation_is_bit_exact(emu, name, count):
    """fast-mode arithmetic == reference DAG on every coefficient of every iteration (tolerance 0)"""
    orc = O.Oracle(name); keys = keys_for(name, with_ksk=(name == "128"))
    bskp = np.empty(orc.bsk_len, np.float64)
    emu.emu_permute_bsk(ptr(keys.bsk), orc.n, orc.L, ptr(bskp))
    bits = np.array([1, 0, 1], np.uint8)
    ca = orc.encrypt_bools(bits, keys, 1); cb = orc.encrypt_bools(bits[::-1].copy(), keys, 2)
    for i in range(count):
        lin = orc.gate_linear(O.XOR if i else O.NAND, ca[i], cb[i])
        ref, tr, m_ref = orc.blind_rotate(lin, keys, trace=True, with_margin=True)
        out = np.empty((2, 1024), np.uint32); tre = np.empty((orc.n, 2, 1024), np.uint32); m = C.c_double(0)
        emu.emu_blind_rotate(orc.n, orc.L, orc.bgbit, C.c_uint32(keys.offset), ptr(lin), ptr(bskp), None, 0, ptr(out), ptr(tre), C.byref(m))
        assert (tre == tr).all(), "first differing iteration %d" % next(j for j in range(orc.n) if (tre[j] != tr[j]).any())
        assert (out == ref).all()
        assert m.value < 0.25 and m_ref < 0.25


def test_emulated_lut_rotation(emu, orc128, keys128):
    """custom test vector (trgsw.zig:336-400) through the emulated kernel arithmetic"""
    bskp = np.empty(orc128.bsk_len, np.float64)
    emu.emu_permute_bsk(ptr(keys128.bsk), orc128.n, orc128.L, ptr(bskp))
    tv = orc128.lut_generate(np.array([2, 0, 3, 1], np.uint32), 4)
    ct = orc128.encrypt_lwe_messages(np.array([3], np.uint32), 4, keys128, seed=9)[0]
    ref = orc128.blind_rotate(ct, keys128, testvec=tv)
    out = np.empty((2, 1024), np.uint32)
    emu.emu_blind_rotate(orc128.n, orc128.L, orc128.bgbit, C.c_uint32(keys128.offset), ptr(ct), ptr(bskp), ptr(tv), 0, ptr(out), None, None)
    assert (out == ref).all()


def test_exact_tables_are_conjugate_and_exact_forward_matches_oracle(emu):
    """exact_fft.cuh: the register-blocked exact transform (three radix-2 stages per pass, bit reversal as register
    naming) returns the reference's forward spectrum bit for bit (the reference's x2 aside, an exact power of two)"""
    assert emu.emu_exact_tables_conjugate() == 1
    rng = np.random.default_rng(3)
    for lo, hi in ((-32, 32), (-2**21, 2**21), (-2**31, 2**31)):
        poly = rng.integers(lo, hi, 1024).astype(np.int32)
        out = np.empty(1024, np.float64)
        emu.emu_exact_forward(ptr(poly), ptr(out))
        assert (out * 2.0 == O.ifft1024(poly.view(np.uint32))).all()


@pytest.mark.parametrize("name,modulus", [("uint4", 16), ("uint1", 2), ("uint2", 4), ("128", 4), ("uint3", 8), ("uint6", 64), ("uint8", 256)])
def test_emulated_exact_blind_rotation_is_bit_exact(emu, name, modulus):
    """exact-mode arithmetic (same __host__ __device__ code as blind_rotate_exact_rb_kernel) == oracle on every coefficient
    of every iteration, on the large-digit sets where the fast transform rounds differently"""
    orc = O.Oracle(name); keys = keys_for(name, with_ksk=(name == "128"))
    bskx = np.empty(orc.bsk_len, np.float64)
    emu.emu_permute_bsk_exact(ptr(keys.bsk), orc.n, orc.L, ptr(bskx))
    msgs = np.array([3 % modulus, 1], np.uint32)
    ct = orc.encrypt_lwe_messages(msgs, modulus, keys, seed=5)
    tv = orc.lut_generate(np.array([(x * x + 1) % modulus for x in range(modulus)], np.uint32), modulus)
    for i in range(2):
        ref, tr = orc.blind_rotate(ct[i], keys, testvec=tv, trace=True)
        out = np.empty((2, 1024), np.uint32); tre = np.empty((orc.n, 2, 1024), np.uint32)
        emu.emu_blind_rotate_exact(orc.n, orc.L, orc.bgbit, C.c_uint32(keys.offset), ptr(ct[i]), ptr(bskx), ptr(tv), ptr(out), ptr(tre))
        assert (tre == tr).all(), "first differing iteration %d" % next(j for j in range(orc.n) if (tre[j] != tr[j]).any())
        assert (out == ref).all()
