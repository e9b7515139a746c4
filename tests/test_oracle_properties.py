"""The reference's own hot-path property tests, re-expressed against the CPU oracle.

These pin the oracle: the reference ships no golden vectors, so its in-file `test "..."` blocks are the
only specification.  Each test names the Zig test it mirrors."""
import numpy as np
import pytest

from conftest import TRUTH, keys_for
from oracle import oracle as O

N = 1024


def _ulps(a, b):
    d = (a.astype(np.int64) - b.astype(np.int64)) & 0xFFFFFFFF
    return np.minimum(d, (1 << 32) - d)


def test_f64_to_torus_constants():
    """utils.zig:28-33 + the gate constants of gates.zig:52-119"""
    assert O.f64_to_torus(0.125) == 0x20000000
    assert O.f64_to_torus(-0.125) == 0xE0000000
    assert O.f64_to_torus(0.25) == 0x40000000
    assert O.f64_to_torus(-0.25) == 0xC0000000
    assert O.f64_to_torus(0.0) == 0
    assert O.f64_to_torus(1.0) == 0 and O.f64_to_torus(-1e-20) == 0     # Zig @mod lowering (see oracle source)
    assert O.f64_to_torus(0.999999999999) == int(0.999999999999 * 2**32)


def test_fft_ifft_roundtrip():
    """fft.zig:783 "fft ifft roundtrip", :873 "fft ifft 1024", :949 "klemsa roundtrip": < 2 torus ulps"""
    rng = np.random.default_rng(42)
    for _ in range(10):
        a = rng.integers(0, 2**32, N, dtype=np.uint32)
        assert _ulps(O.fft1024(O.ifft1024(a)), a).max() < 2


def test_fft_poly_mul_vs_schoolbook():
    """fft.zig:814 "fft poly mul", :914 "fft poly mul 1024": uniform a x (b < BG) within < 2 ulps"""
    rng = np.random.default_rng(42)
    for _ in range(20):
        a = rng.integers(0, 2**32, N, dtype=np.uint32)
        b = rng.integers(0, 64, N, dtype=np.uint32)
        assert _ulps(O.poly_mul_fft(a, b), O.poly_mul_naive(a, b)).max() < 2


def test_generic_radix2_small():
    """fft.zig:725 "simple fft test", :752 "delta function test" (N = 8 generic path)"""
    x = np.arange(1, 9, dtype=np.float64) + 0j
    assert np.allclose(O.radix2_fft(x), np.fft.fft(x), atol=1e-12)
    assert np.allclose(O.radix2_fft(O.radix2_fft(x), inverse=True) / 8, x, atol=1e-12)
    delta = np.zeros(8, complex); delta[0] = 1
    assert np.allclose(O.radix2_fft(delta), np.ones(8), atol=1e-15)


def test_poly_mul_with_xk():
    """trgsw.zig:757 "trgsw poly mul with x k": k=0 identity, k=1 wraps -a[N-1], k=N negation, k=2N identity"""
    a = np.arange(1, N + 1, dtype=np.uint32)
    assert (O.poly_mul_with_xk(a, 0) == a).all()
    r1 = O.poly_mul_with_xk(a, 1)
    assert r1[0] == (0 - int(a[N - 1])) & 0xFFFFFFFF and (r1[1:] == a[:-1]).all()
    assert (O.poly_mul_with_xk(a, N) == (0 - a.astype(np.int64)).astype(np.uint32)).all()
    assert (O.poly_mul_with_xk(a, 2 * N) == a).all()
    # against the in-repo spec of negacyclic multiplication (fft.zig:695-714) for a few k
    for k in (3, 511, 1023, 1025, 2047):
        xk = np.zeros(N, np.uint32)
        if k < N:
            xk[k] = 1
        else:
            xk[k - N] = 0xFFFFFFFF
        assert (O.poly_mul_with_xk(a, k) == O.poly_mul_naive(a, xk)).all()


def test_sample_extraction(orc128):
    """trlwe.zig:296 "sample extraction deterministic" + :229 "sample extract index" """
    rng = np.random.default_rng(1)
    t = rng.integers(0, 2**32, (2, N), dtype=np.uint32)
    for k in (0, 1, 17, N - 1):
        s = orc128.sample_extract_index(t, k)
        assert s[N] == t[1, k] and s[0] == t[0, k]
    s0 = orc128.sample_extract_index(t, 0)
    assert (s0[1:N] == (0 - t[0, :0:-1].astype(np.int64)).astype(np.uint32)).all()
    keys = keys_for("128")
    bits = rng.integers(0, 2, N).astype(np.uint8)
    c = orc128.trlwe_encrypt_bools(bits, keys, seed=3)
    for k in (0, 5, 1000):
        lv1 = orc128.sample_extract_index(c, k)
        assert orc128.decrypt_bools(lv1, keys, level=1)[0] == bits[k]


def test_decomposition_reconstructs(orc128, keys128):
    """trgsw.zig:505 "trgsw decomposition": sum digit_k * BG^-(k+1) decrypts to the plaintext"""
    rng = np.random.default_rng(42)
    h = [O.f64_to_torus(64.0 ** -(i + 1)) for i in range(3)]
    for trial in range(3):
        bits = rng.integers(0, 2, N).astype(np.uint8)
        c = orc128.trlwe_encrypt_bools(bits, keys128, seed=10 + trial)
        dec = orc128.decomposition(c, keys128.offset)
        assert (dec.view(np.int32) >= -32).all() and (dec.view(np.int32) <= 31).all()
        rec = np.zeros((2, N), np.uint32)
        for k in range(3):
            rec[0] += dec[k] * np.uint32(h[k])
            rec[1] += dec[k + 3] * np.uint32(h[k])
        assert (orc128.trlwe_decrypt_bools(rec, keys128) == bits).all()


def test_external_product_and_cmux(orc128, keys128):
    """trgsw.zig:578 "trgsw external product with fft", :637 "trgsw cmux" """
    rng = np.random.default_rng(42)
    one = orc128.trgsw_encrypt_fft(1, keys128, seed=5)
    zero = orc128.trgsw_encrypt_fft(0, keys128, seed=6)
    for trial in range(2):
        b1 = rng.integers(0, 2, N).astype(np.uint8); b2 = rng.integers(0, 2, N).astype(np.uint8)
        c1 = orc128.trlwe_encrypt_bools(b1, keys128, seed=20 + trial)
        c2 = orc128.trlwe_encrypt_bools(b2, keys128, seed=30 + trial)
        ep = orc128.external_product(one, c1, keys128.offset)
        assert (orc128.trlwe_decrypt_bools(ep, keys128) == b1).all()
        assert (orc128.trlwe_decrypt_bools(orc128.cmux(c1, c2, zero, keys128.offset), keys128) == b1).all()
        assert (orc128.trlwe_decrypt_bools(orc128.cmux(c1, c2, one, keys128.offset), keys128) == b2).all()


def test_integer_oracle_agrees_bit_for_bit(orc128, keys128):
    """SURVEY.md section 8c second oracle: on L=3/BGBIT=6 sets the FP64 external product equals the exact
    integer negacyclic convolution, with a wide rounding margin."""
    rng = np.random.default_rng(7)
    for i in (0, 123, 699):
        t = rng.integers(0, 2**32, (2, N), dtype=np.uint32)
        fp, margin = orc128.external_product(keys128.bsk[i], t, keys128.offset, with_margin=True)
        assert (fp == orc128.external_product_int(keys128.bsk[i], t, keys128.offset)).all()
        assert margin < 0.25


def test_integer_oracle_follows_a_whole_blind_rotation(orc128, keys128):
    """all n rows of one bootstrap: the FP64 oracle's accumulator after every CMUX step equals
    acc + (exact integer negacyclic convolution of the digits of X^a_i acc - acc with row i of the key), mod 2^32 --
    i.e. the reference's f64 transform path never rounds away from the integer result on this trajectory"""
    ct = orc128.encrypt_bools(np.array([1], np.uint8), keys128, seed=3)[0]
    out, trace, margin = orc128.blind_rotate(ct, keys128, trace=True, with_margin=True)
    assert margin < 0.25
    btil = 2 * N - ((int(ct[-1]) + (1 << 20)) >> 21)                          # trgsw.zig:297
    acc = np.stack([O.poly_mul_with_xk(keys128.testvec[h], btil) for h in range(2)])
    for i in range(orc128.n):
        at = (int(ct[i]) + (1 << 20)) >> 21                                     # trgsw.zig:312
        rot = np.stack([O.poly_mul_with_xk(acc[h], at) for h in range(2)])
        acc = acc + orc128.external_product_int(keys128.bsk[i], rot - acc, keys128.offset)   # cmux, trgsw.zig:260-284
        assert (acc == trace[i]).all(), f"row {i}"
    assert (acc == out).all()


def test_blind_rotate_and_key_switch_decrypt(orc128, keys128):
    """trgsw.zig:694 "trgsw blind rotate" (we demand 100 %, the reference only 60 %), :729 "identity key switching" """
    bits = np.array([0, 1, 1, 0, 1], np.uint8)
    ct = orc128.encrypt_bools(bits, keys128, seed=77)
    for i, b in enumerate(bits):
        tr, margin = orc128.blind_rotate(ct[i], keys128, with_margin=True)
        assert margin < 0.25
        lv1 = orc128.sample_extract_index(tr, 0)
        assert orc128.decrypt_bools(lv1, keys128, level=1)[0] == b
        lv0 = orc128.identity_key_switching(lv1, keys128)
        assert orc128.decrypt_bools(lv0, keys128)[0] == b
        # key switching is linear: number of touched rows ~ N*t*(base-1)/base
    lv1 = np.zeros(N + 1, np.uint32); lv1[N] = 12345
    assert orc128.identity_key_switching(lv1, keys128)[orc128.n] != 0


@pytest.mark.parametrize("op", [O.NAND, O.AND, O.OR, O.XOR, O.NOR, O.XNOR, O.ANDNY, O.ANDYN, O.ORNY, O.ORYN])
def test_gate_truth_tables(orc128, keys128, op):
    """gates.zig:374-511 "gates all NAND/AND/OR/XOR/NOR cases" (the other five gates are untested upstream;
    XNOR decrypts as XOR under the reference's a - 2b - 1/4, gates.zig:78-82 -- kept)."""
    a = np.array([0, 0, 1, 1], np.uint8); b = np.array([0, 1, 0, 1], np.uint8)
    ca = orc128.encrypt_bools(a, keys128, 101); cb = orc128.encrypt_bools(b, keys128, 102)
    out = orc128.gate_batch(op, ca, cb, keys128)
    assert (orc128.decrypt_bools(out, keys128) == TRUTH[op](a, b)).all()


def test_mux_not_constant(orc128, keys128):
    """gates.zig:513 "gates mux naive", :131-151 NOT / CONSTANT (false = 1 - 2^29 quirk)"""
    for a in (0, 1):
        for b in (0, 1):
            for c in (0, 1):
                ca, cb, cc = (orc128.encrypt_bools(np.array([x], np.uint8), keys128, 200 + 4 * a + 2 * b + c + 10 * i)[0]
                              for i, x in enumerate((a, b, c)))
                a_and_b = orc128.gate(O.AND, ca, cb, keys128)
                nand_a_c = orc128.gate(O.AND, orc128.gate_not(ca), cc, keys128)
                out = orc128.gate(O.OR, a_and_b, nand_a_c, keys128)
                assert orc128.decrypt_bools(out, keys128)[0] == (b if a else c)
    assert orc128.gate_constant(True)[-1] == 0x20000000
    assert orc128.gate_constant(False)[-1] == 0xE0000001
    assert orc128.decrypt_bools(orc128.gate_constant(True), keys128)[0] == 1
    assert orc128.decrypt_bools(orc128.gate_constant(False), keys128)[0] == 0


def test_lut_generator_and_encoder(orc128):
    """lut/encoder.zig:66-105, lut/generator.zig:85-135 (the reference's generator tests pin nothing but sizes)"""
    for m in (2, 4, 8, 16):
        for x in range(m):
            assert O.lut_decode(O.lut_encode(x, m), m) == x
    m = 4
    tv = orc128.lut_generate(np.array([0, 1, 2, 3], np.uint32), m)
    assert (tv[0] == 0).all()
    off = (N + m) // (2 * m)            # divRound(N, 2m)
    assert tv[1][0] == O.lut_encode(0, m) and tv[1][N // m] == O.lut_encode(1, m)
    assert tv[1][N - 1] == (0 - O.lut_encode(0, m)) & 0xFFFFFFFF and off == 128


def test_lut_bootstrap_functional_128(orc128, keys128):
    """programmable bootstrap f(m) at the 128-bit set, m = 4 (UINT3+ sets cannot decode)"""
    m = 4
    msgs = np.arange(4, dtype=np.uint32)
    ct = orc128.encrypt_lwe_messages(msgs, m, keys128, seed=55)
    assert (orc128.decrypt_lwe_messages(ct, m, keys128) == msgs).all()
    table = np.array([1, 3, 0, 2], np.uint32)
    out = orc128.bootstrap_batch(ct, keys128, orc128.lut_generate(table, m))
    assert (orc128.decrypt_lwe_messages(out, m, keys128) == table[msgs]).all()


@pytest.mark.parametrize("name", ["80", "110"])
def test_other_sets_decrypt(name):
    orc = O.Oracle(name); keys = keys_for(name)
    a = np.array([0, 1], np.uint8); b = np.array([1, 1], np.uint8)
    out = orc.gate_batch(O.NAND, orc.encrypt_bools(a, keys, 1), orc.encrypt_bools(b, keys, 2), keys)
    assert (orc.decrypt_bools(out, keys) == 1 - (a & b)).all()


def test_key_material_shapes_and_constants(orc128, keys128):
    """key.zig:214-276 "secret key generation", "decomposition offset generation", "test vector generation",
    "key switching key generation"; trlwe.zig:273 "trlwe fft representation" """
    assert set(np.unique(keys128.s0)) <= {0, 1} and set(np.unique(keys128.s1)) <= {0, 1}
    assert keys128.s0.any() and keys128.s1.any()
    assert keys128.offset == 0x82080000 != 0                        # sum_i 32 * 2^(32 - 6(i+1)), key.zig:121-131
    assert (keys128.testvec[0] == 0).all() and (keys128.testvec[1] == 0x20000000).all()
    assert keys128.ksk.shape == (1024 * 9 * 4, 701) and keys128.bsk.shape == (700, 6, 2, 1024)
    assert (keys128.ksk.reshape(1024, 9, 4, 701)[:, :, 0, :] == 0).all()   # k = 0 rows never written (key.zig:159)
    # a BSK row is TRGSW(s0_i) in FFT form: transforming back gives integers (exact round trip)
    back, m = O.fft1024(keys128.bsk[3, 0, 0], with_margin=True)
    assert m < 1e-3 and (O.ifft1024(back) == keys128.bsk[3, 0, 0]).all()
    for name, off in (("80", 0x82080000), ("uint1", 0x80200000), ("uint4", 0x80000000)):
        assert O.Oracle(name).keygen.__self__ is not None
        import ctypes as C
        assert int(O.lib().orc_decomposition_offset(C.byref(O.Oracle(name).p))) == off


def test_tlwe_lwe_message_encoding(orc128, keys128):
    """tlwe.zig:370 "tlwe lwe message encoding" (the reference only demands 8 of 10)"""
    ct = orc128.encrypt_lwe_messages(np.full(10, 2, np.uint32), 4, keys128, seed=123)
    assert (orc128.decrypt_lwe_messages(ct, 4, keys128) == 2).all()
    bits = np.array([1, 0, 1, 1, 0], np.uint8)
    assert (orc128.decrypt_bools(orc128.encrypt_bools(bits, keys128, 9), keys128) == bits).all()   # tlwe.zig:300
