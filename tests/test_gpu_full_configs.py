"""BASELINE.json configs at FULL size, bit-compared with the CPU oracle on wide samples.

Every test here calls the CUDA path through the C ABI at the batch size the config names and compares
ciphertext words (not only decrypted bits) with the oracle on >= 256 ... >= 1,024 items chosen to cover the
CTA tile (4 ciphertexts), the SM wave (148 x 4 = 592), the tail launch and the per-launch chunk (2^18).
The oracle costs ~2 ms per gate per host thread-second, so a thousand items is a few seconds."""
import numpy as np
import pytest

from conftest import TRUTH, keys_for
from oracle import oracle as O

pytestmark = pytest.mark.gpu

WAVE = 148 * 4


def _ctx(name, mode=None):
    import tfhe_b200
    k = keys_for(name)
    c = tfhe_b200.Context(name, devices=[0])
    c.load_key(k.bsk, k.ksk, k.offset)
    if mode is not None:
        c.set_mode(mode)
    return c, O.Oracle(name), k


def _enc_pairs(orc, keys, B, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 2, B).astype(np.uint8)
    b = rng.integers(0, 2, B).astype(np.uint8)
    return a, b, orc.encrypt_bools(a, keys, seed * 2 + 1), orc.encrypt_bools(b, keys, seed * 2 + 2)


def _boundary_sample(B, count, seed, extra=()):
    """`count` distinct indices of [0, B): both ends, every multiple of the wave / tile next to a boundary, `extra`, then a
    uniform random fill"""
    rng = np.random.default_rng(seed)
    pts = {0, 1, 2, 3, 4, 5, B - 1, B - 2, B - 3, B - 4, B - 5}
    last_full = (B // WAVE) * WAVE
    for base in (WAVE, 2 * WAVE, 55 * WAVE, last_full, B // 2):
        pts.update(range(base - 5, base + 5))
    pts.update(extra)
    pts = {p for p in pts if 0 <= p < B}
    fill = rng.permutation(B)
    for p in fill:
        if len(pts) >= count:
            break
        pts.add(int(p))
    return np.array(sorted(pts), np.int64)


def test_config2_65536_and_xor_1024_items_vs_oracle():
    """BASELINE config 2: 65,536 independent AND/XOR gates at the 128-bit set; 1,100+ items -- ends, wave and tail
    boundaries (the launch is 110 whole waves + a 416-ciphertext tail), the AND/XOR switch, random fill -- word for word."""
    c, orc, k = _ctx("128")
    try:
        B = 65536
        a, b, ca, cb = _enc_pairs(orc, k, B, seed=42)
        ops = np.where(np.arange(B) < B // 2, O.AND, O.XOR).astype(np.int32)
        got = c.gate_batch(ops, ca, cb)
        truth = np.where(ops == O.AND, a & b, a ^ b)
        assert (orc.decrypt_bools(got, k) == truth).all()
        sel = _boundary_sample(B, 1100, seed=1)
        assert len(sel) >= 1024
        ref = orc.gate_batch(ops[sel], ca[sel], cb[sel], k)
        bad = (got[sel] != ref).any(axis=1)
        assert not bad.any(), f"{bad.sum()} of {len(sel)} sampled gates differ, first at {sel[bad][:5]}"
        # the host path: pageable buffers above (four pipelined chunks, staged copies); the same batch from PINNED buffers
        # (pipelined too, plain asynchronous copies) and as one launch pair (host_pipeline = 0) must give the same words
        import torch
        pa = torch.from_numpy(ca).pin_memory(); pb = torch.from_numpy(cb).pin_memory()
        po = torch.empty((B, ca.shape[1]), dtype=torch.int32).pin_memory()
        out_pinned = po.numpy().view(np.uint32)
        c.gate_batch(ops, pa.numpy(), pb.numpy(), out=out_pinned)
        assert (out_pinned == got).all()
        c.set_tuning("host_pipeline", 0)
        assert (c.gate_batch(ops, ca, cb) == got).all()
    finally:
        c.close()


def test_config3_adder_1024_instances_sample_vs_oracle():
    """BASELINE config 3: 1,024 x 16-bit ripple-carry additions through the native circuit executor (4 lanes, CUDA graph);
    32 instances spread over the lanes are re-evaluated gate by gate by the oracle in the reference's order
    (examples/add_two_numbers.zig:24-73) and every output ciphertext word is compared"""
    import tfhe_b200
    from tfhe_b200 import circuits
    c, orc, k = _ctx("128")
    try:
        rng = np.random.default_rng(7)
        W, B = 16, 1024
        x = rng.integers(0, 2**16, B); y = rng.integers(0, 2**16, B)
        x[0], y[0] = 402, 304
        enc = lambda bits, seed: np.stack([orc.encrypt_bools(bits[i], k, seed + i) for i in range(W)])
        ca, cb = enc(circuits.to_bits(x, W), 1000), enc(circuits.to_bits(y, W), 2000)
        cin = orc.encrypt_bools(np.zeros(B, np.uint8), k, 3000)
        sums, carry, circ = circuits.ripple_carry_add_native(c, ca, cb, cin)
        dec = np.stack([orc.decrypt_bools(sums[i], k) for i in range(W)])
        total = circuits.from_bits(dec) + (orc.decrypt_bools(carry, k).astype(np.uint64) << np.uint64(W))
        assert (total == x + y).all() and total[0] == 706
        sel = np.unique(np.concatenate([[0, 1, 255, 256, 257, 511, 512, 767, 768, 1022, 1023], rng.integers(0, B, 24)]))[:32]
        cr = cin[sel]
        for i in range(W):                                   # fullAdder, bit by bit, batched over the sampled instances
            axb = orc.gate_batch(O.XOR, ca[i, sel], cb[i, sel], k); ab = orc.gate_batch(O.AND, ca[i, sel], cb[i, sel], k)
            t = orc.gate_batch(O.AND, axb, cr, k); s = orc.gate_batch(O.XOR, axb, cr, k)
            cr = orc.gate_batch(O.OR, ab, t, k)
            assert (sums[i, sel] == s).all(), f"sum bit {i}"
        assert (carry[sel] == cr).all()
        circ.close()
    finally:
        c.close()


def test_config4_uint4_lut_32768_exact_mode_sample_vs_oracle():
    """BASELINE config 4: 32,768 programmable bootstraps at SECURITY_UINT4 in exact mode (the mode in which this set equals
    the oracle); three function tables, per-item; 320 items word for word, before and after the key switch"""
    import tfhe_b200
    c, orc, k = _ctx("uint4", tfhe_b200.MODE_EXACT)
    try:
        B, m = 32768, 16
        rng = np.random.default_rng(3)
        msgs = rng.integers(0, m, B).astype(np.uint32)
        ct = orc.encrypt_lwe_messages(msgs, m, k, seed=5)
        tables = [np.arange(m, dtype=np.uint32), (np.arange(m, dtype=np.uint32) + 1) % m, (np.arange(m, dtype=np.uint32) ** 2) % m]
        tvs = [orc.lut_generate(t, m) for t in tables]
        got = c.bootstrap_batch(ct, tvs[2])
        sel = _boundary_sample(B, 320, seed=2)
        assert len(sel) >= 256
        ref = orc.bootstrap_batch(ct[sel], k, tvs[2])
        bad = (got[sel] != ref).any(axis=1)
        assert not bad.any(), f"{bad.sum()} of {len(sel)} sampled LUT bootstraps differ, first at {sel[bad][:5]}"
        # pre-keyswitch accumulators of a smaller sample, other tables
        sub = sel[::10]
        for tv in tvs[:2]:
            assert (c.blind_rotate_batch(ct[sub], tv) == orc.blind_rotate_batch(ct[sub], k, tv)).all()
    finally:
        c.close()


@pytest.mark.parametrize("name", ["80", "110"])
def test_config5_other_sets_mixed_opcodes_two_waves(name):
    """BASELINE config 5's other parameter sets: all ten opcodes mixed over more than two CTA waves (1,500 gates = 2 waves +
    a 316-ciphertext tail), every decrypted bit against the truth table, 300 items word for word"""
    c, orc, k = _ctx(name)
    try:
        B = 1500
        a, b, ca, cb = _enc_pairs(orc, k, B, seed=17)
        ops = (np.arange(B) * 7 % 10).astype(np.int32)
        got = c.gate_batch(ops, ca, cb)
        want = np.array([TRUTH[int(ops[i])](int(a[i]), int(b[i])) for i in range(B)], np.uint8)
        assert (orc.decrypt_bools(got, k) == want).all()
        sel = _boundary_sample(B, 300, seed=4)
        ref = orc.gate_batch(ops[sel], ca[sel], cb[sel], k)
        assert (got[sel] == ref).all()
    finally:
        c.close()


def test_config5_chunk_boundary_2pow18_bit_compared():
    """one host call larger than the per-launch chunk (2^18 ciphertexts): the items either side of the chunk boundary and of
    the second chunk's tail are compared word for word; every bit is decrypted"""
    c, orc, k = _ctx("128")
    try:
        B = (1 << 18) + 700
        a, b, ca, cb = _enc_pairs(orc, k, B, seed=23)
        got = c.gate_batch(O.NAND, ca, cb)
        assert (orc.decrypt_bools(got, k) == 1 - (a & b)).all()
        edge = 1 << 18
        sel = _boundary_sample(B, 256, seed=6, extra=range(edge - 12, edge + 12))
        ref = orc.gate_batch(O.NAND, ca[sel], cb[sel], k)
        assert (got[sel] == ref).all()
    finally:
        c.close()
