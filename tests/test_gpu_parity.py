"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on identical
keys and ciphertexts.  Bit-exact everywhere (tolerance 0) on the L=3/BGBIT=6 sets: integer steps are
exactly specified, and the FP64 external product is an exact integer computation there
(the margin counter proves it stays true on the device)."""
import numpy as np
import pytest

from conftest import TRUTH, keys_for
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx128():
    import tfhe_b200
    k = keys_for("128")
    c = tfhe_b200.Context("128", devices=[0])
    c.load_key(k.bsk, k.ksk, k.offset)
    yield c
    c.close()


def _enc_pairs(orc, keys, B, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 2, B).astype(np.uint8)
    b = rng.integers(0, 2, B).astype(np.uint8)
    return a, b, orc.encrypt_bools(a, keys, seed * 2 + 1), orc.encrypt_bools(b, keys, seed * 2 + 2)


def test_library_is_cuda_native(ctx128):
    assert ctx128.launch_count() >= 1   # key re-layout kernels already ran
    assert ctx128.stream(0) != 0


def test_keyswitch_bit_exact(ctx128, orc128, keys128):
    rng = np.random.default_rng(5)
    for B in (1, 7, 40, 700):
        lv1 = rng.integers(0, 2**32, (B, 1025), dtype=np.uint32)
        got = ctx128.keyswitch_batch(lv1)
        ref = orc128.keyswitch_batch(lv1, keys128)
        assert (got == ref).all(), f"key switch differs at B={B}"


@pytest.mark.parametrize("use_tma,kct,twt", [(1, 0, 0), (0, 0, 0), (1, 1, 0), (1, 2, 0), (1, 3, 0), (1, 4, 0), (1, 5, 0), (1, 6, 0), (0, 6, 0),
                                             (1, 6, -1), (1, 5, -1), (1, 4, 1), (1, 5, 1)])
def test_blind_rotate_bit_exact(ctx128, orc128, keys128, use_tma, kct, twt):
    """pre-keyswitch TRLWE coefficients: stated tolerance 0 versus the reference f64 FFT path.  Every CTA width of the
    throughput kernel, key ring or direct loads, twiddles in registers / expanded / in tensor memory (twt), accumulators
    in registers or tensor memory (kct 6, twt -1)."""
    ctx128.set_tuning("use_tma", use_tma)
    ctx128.set_tuning("kct", kct)
    ctx128.set_tuning("twt", twt)
    try:
        B = 13
        _, _, ca, cb = _enc_pairs(orc128, keys128, B, seed=3)
        lin = np.stack([orc128.gate_linear(O.NAND, ca[i], cb[i]) for i in range(B)])
        got = ctx128.blind_rotate_batch(lin)
        ref = orc128.blind_rotate_batch(lin, keys128)
        assert (got == ref).all(), f"{(got != ref).any(axis=(1, 2)).sum()} of {B} accumulators differ"
        lv1 = ctx128.blind_rotate_extract_batch(lin)
        ref_lv1 = np.stack([orc128.sample_extract_index(ref[i], 0) for i in range(B)])
        assert (lv1 == ref_lv1).all()
    finally:
        ctx128.set_tuning("use_tma", 1)
        ctx128.set_tuning("kct", 0)
        ctx128.set_tuning("twt", 0)


@pytest.mark.parametrize("op", list(range(10)))
def test_gate_truth_tables(ctx128, orc128, keys128, op):
    """gates.zig:374-544 truth tables through real bootstraps + bit-exact ciphertext parity."""
    a = np.array([0, 0, 1, 1], np.uint8); b = np.array([0, 1, 0, 1], np.uint8)
    ca = orc128.encrypt_bools(a, keys128, 21); cb = orc128.encrypt_bools(b, keys128, 22)
    got = ctx128.gate_batch(op, ca, cb)
    ref = orc128.gate_batch(op, ca, cb, keys128)
    assert (orc128.decrypt_bools(got, keys128) == TRUTH[op](a, b)).all()
    assert (got == ref).all()


def test_mixed_and_xor_batch(ctx128, orc128, keys128):
    B = 300
    a, b, ca, cb = _enc_pairs(orc128, keys128, B, seed=9)
    ops = np.where(np.arange(B) % 2 == 0, O.AND, O.XOR).astype(np.int32)
    got = ctx128.gate_batch(ops, ca, cb)
    truth = np.where(ops == O.AND, a & b, a ^ b)
    assert (orc128.decrypt_bools(got, keys128) == truth).all()
    sel = np.arange(0, B, 7)
    ref = orc128.gate_batch(ops[sel], ca[sel], cb[sel], keys128)
    assert (got[sel] == ref).all()


def test_round_margin_counter(ctx128, orc128, keys128):
    """standing proof of exactness: max |t - round(t)| in the inverse-transform epilogue << 0.5"""
    _, _, ca, cb = _enc_pairs(orc128, keys128, 64, seed=4)
    ctx128.track_margin(True)
    try:
        ctx128.max_round_margin(reset=True)
        got = ctx128.gate_batch(O.NAND, ca, cb)
        m = ctx128.max_round_margin(reset=True)
    finally:
        ctx128.track_margin(False)
    assert 0.0 < m < 0.25, m
    assert (got == ctx128.gate_batch(O.NAND, ca, cb)).all()   # margin variant == production variant


def test_bootstrap_without_keyswitch_and_not(ctx128, orc128, keys128):
    a, _, ca, _ = _enc_pairs(orc128, keys128, 5, seed=6)
    got = ctx128.bootstrap_no_keyswitch_batch(ca)
    for i in range(5):
        tr = orc128.blind_rotate(ca[i], keys128)
        assert (got[i] == orc128.sample_extract_index2(tr, 0)).all()
    assert (ctx128.not_batch(ca) == (0 - ca.astype(np.int64)).astype(np.uint32)).all()
    assert (ctx128.bootstrap_batch(ca) == orc128.bootstrap_batch(ca, keys128)).all()


def test_lut_bootstrap_128(ctx128, orc128, keys128):
    """programmable bootstrap (trgsw.zig:336-400 + lut/generator.zig:85-135) at the 128-bit set, m = 4."""
    m = 4
    msgs = np.arange(16, dtype=np.uint32) % m
    ct = orc128.encrypt_lwe_messages(msgs, m, keys128, seed=31)
    for table in ([0, 1, 2, 3], [1, 2, 3, 0], [0, 1, 0, 1]):
        tv = orc128.lut_generate(np.array(table, np.uint32), m)
        got = ctx128.bootstrap_batch(ct, tv)
        ref = orc128.bootstrap_batch(ct, keys128, tv)
        assert (got == ref).all()
        dec = orc128.decrypt_lwe_messages(got, m, keys128)
        assert (dec == np.array(table, np.uint32)[msgs]).all(), (table, dec)
    # per-item test vectors
    tvs = np.stack([orc128.lut_generate(np.array([(x + s) % m for x in range(m)], np.uint32), m) for s in range(16)])
    got = ctx128.bootstrap_batch(ct, tvs, tv_per_item=True)
    ref = orc128.bootstrap_batch(ct, keys128, tvs, tv_per_item=True)
    assert (got == ref).all()


@pytest.mark.parametrize("name", ["80", "110"])
def test_other_security_levels_bit_exact(name):
    import tfhe_b200
    orc = O.Oracle(name); keys = keys_for(name)
    c = tfhe_b200.Context(name, devices=[0])
    try:
        c.load_key(keys.bsk, keys.ksk, keys.offset)
        a, b, ca, cb = _enc_pairs(orc, keys, 24, seed=8)
        got = c.gate_batch(O.NAND, ca, cb)
        assert (orc.decrypt_bools(got, keys) == 1 - (a & b)).all()
        assert (got == orc.gate_batch(O.NAND, ca, cb, keys)).all()
    finally:
        c.close()


def test_full_size_batch_properties(ctx128, orc128, keys128):
    """BASELINE config 2 at full size (65,536 AND/XOR): every decrypted bit equals the plaintext truth,
    a strided sample is bit-exact against the oracle, and the result is independent of the tiling."""
    B = 65536
    a, b, ca, cb = _enc_pairs(orc128, keys128, B, seed=42)
    ops = np.where(np.arange(B) < B // 2, O.AND, O.XOR).astype(np.int32)
    got = ctx128.gate_batch(ops, ca, cb)
    truth = np.where(ops == O.AND, a & b, a ^ b)
    assert (orc128.decrypt_bools(got, keys128) == truth).all()
    sel = np.arange(0, B, 2048)
    assert (got[sel] == orc128.gate_batch(ops[sel], ca[sel], cb[sel], keys128)).all()
    ctx128.set_tuning("kct", 4)
    try:
        again = ctx128.gate_batch(ops[:1000], ca[:1000], cb[:1000])
    finally:
        ctx128.set_tuning("kct", 0)
    assert (again == got[:1000]).all()


def test_errors_are_loud(orc128, keys128):
    import tfhe_b200
    c = tfhe_b200.Context("128", devices=[0])
    try:
        with pytest.raises(tfhe_b200.TfheB200Error):
            c.gate_batch(O.NAND, np.zeros((1, 701), np.uint32), np.zeros((1, 701), np.uint32))   # no key
        k = keys_for("128")
        c.load_key(k.bsk, None, k.offset)       # CloudKey.newNoKsk analogue
        c.blind_rotate_batch(np.zeros((1, 701), np.uint32))
        with pytest.raises(tfhe_b200.TfheB200Error):
            c.gate_batch(O.NAND, np.zeros((1, 701), np.uint32), np.zeros((1, 701), np.uint32))   # no KSK
        with pytest.raises(tfhe_b200.TfheB200Error):
            c.gate_batch(17, np.zeros((1, 701), np.uint32), np.zeros((1, 701), np.uint32))
    finally:
        c.close()


def test_edge_cases_empty_ragged_and_wave_boundaries(ctx128, orc128, keys128):
    """empty batch, batch sizes straddling the CTA tile (4 ciphertexts) and the SM-wave boundaries, gate == bootstrap(linear)"""
    w = 701
    assert ctx128.gate_batch(O.NAND, np.zeros((0, w), np.uint32), np.zeros((0, w), np.uint32)).shape == (0, w)
    assert ctx128.keyswitch_batch(np.zeros((0, 1025), np.uint32)).shape == (0, w)
    a, b, ca, cb = _enc_pairs(orc128, keys128, 1190, seed=12)
    full = ctx128.gate_batch(O.OR, ca, cb)
    assert (orc128.decrypt_bools(full, keys128) == (a | b)).all()
    for B in (1, 2, 3, 5, 147, 149, 593):          # kct policy switches at 148, 296, 444, 592
        part = ctx128.gate_batch(O.OR, ca[:B], cb[:B])
        assert (part == full[:B]).all(), B
    lin = np.stack([orc128.gate_linear(O.OR, ca[i], cb[i]) for i in range(3)])
    assert (ctx128.bootstrap_batch(lin) == full[:3]).all()
    # extreme inputs: all-zero and all-ones ciphertexts (atil = 0 and 2N edge of the modulus switch)
    z = np.zeros((2, w), np.uint32); z[1] = 0xFFFFFFFF
    assert (ctx128.bootstrap_batch(z) == orc128.bootstrap_batch(z, keys128)).all()
    assert (ctx128.blind_rotate_batch(z) == orc128.blind_rotate_batch(z, keys128)).all()


def test_chunked_launches_equal_single_launch(ctx128, orc128, keys128):
    a, b, ca, cb = _enc_pairs(orc128, keys128, 700, seed=13)
    one = ctx128.gate_batch(O.AND, ca, cb)
    ctx128.set_tuning("max_chunk", 256)
    try:
        assert (ctx128.gate_batch(O.AND, ca, cb) == one).all()
    finally:
        ctx128.set_tuning("max_chunk", 1 << 18)


def test_latency_mode_equals_throughput_mode(ctx128, orc128, keys128):
    """batches <= SM count / 2 run one two-CTA cluster per ciphertext (each CTA owns one accumulator half, partial sums
    cross through distributed shared memory), batches <= SM count one CTA per ciphertext with the 2L transforms in
    parallel; same bits as the throughput kernel"""
    a, b, ca, cb = _enc_pairs(orc128, keys128, 100, seed=14)
    pair = ctx128.gate_batch(O.XOR, ca[:40], cb[:40])        # 40 <= 74: cluster kernel
    lat = ctx128.gate_batch(O.XOR, ca, cb)                   # 74 < 100 <= 148: single-CTA latency kernel
    ctx128.set_tuning("latency_mode", 2)
    try:
        lat40 = ctx128.gate_batch(O.XOR, ca[:40], cb[:40])   # single-CTA latency kernel forced
        ctx128.set_tuning("latency_mode", 0)
        thr = ctx128.gate_batch(O.XOR, ca, cb)
    finally:
        ctx128.set_tuning("latency_mode", 1)
    assert (lat == thr).all() and (pair == thr[:40]).all() and (lat40 == thr[:40]).all()
    one = ctx128.gate_batch(O.XOR, ca[:1], cb[:1])
    assert (one == thr[:1]).all()
    tr = ctx128.blind_rotate_batch(np.stack([orc128.gate_linear(O.NAND, ca[i], cb[i]) for i in range(5)]))   # TRLWE output path of the pair kernel
    assert (tr == orc128.blind_rotate_batch(np.stack([orc128.gate_linear(O.NAND, ca[i], cb[i]) for i in range(5)]), keys128)).all()
    assert (lat[:6] == orc128.gate_batch(O.XOR, ca[:6], cb[:6], keys128)).all()
    ctx128.track_margin(True)
    try:
        ctx128.max_round_margin(reset=True)
        ctx128.gate_batch(O.XOR, ca, cb)
        assert 0.0 < ctx128.max_round_margin(reset=True) < 0.25
    finally:
        ctx128.track_margin(False)


def test_split_launch_full_waves_plus_tail(ctx128, orc128, keys128):
    """a batch that is not a whole number of CTA waves runs as (full waves at KCT = 4) + (tail launch with its own
    kernel: here 5 ciphertexts on the cluster latency kernel) over global ciphertext indices; per-item opcodes and
    per-item test vectors follow the indices"""
    B = 148 * 4 + 5
    a, b, ca, cb = _enc_pairs(orc128, keys128, B, seed=21)
    ops = (np.arange(B) % 10).astype(np.int32)
    split = ctx128.gate_batch(ops, ca, cb)
    ctx128.set_tuning("kct", 4)                      # explicit width: one launch
    try:
        one = ctx128.gate_batch(ops, ca, cb)
    finally:
        ctx128.set_tuning("kct", 0)
    assert (split == one).all()
    from conftest import TRUTH
    want = np.array([TRUTH[int(ops[i])](int(a[i]), int(b[i])) for i in range(B)], np.uint8)
    assert (orc128.decrypt_bools(split, keys128) == want).all()
    tail = slice(B - 7, B)
    ref = np.stack([orc128.gate(int(ops[i]), ca[i], cb[i], keys128) for i in range(B - 7, B)])
    assert (split[tail] == ref).all()
    # per-item test vectors across the split (programmable bootstrap, trgsw.zig:336-400)
    rng = np.random.default_rng(5)
    tv = rng.integers(0, 2**32, (B, 2, 1024), dtype=np.uint32)
    got = ctx128.blind_rotate_batch(ca, tv, tv_per_item=True)
    for i in (0, 591, 592, B - 1):
        assert (got[i] == orc128.blind_rotate_batch(ca[i:i + 1], keys128, tv[i])[0]).all(), i


@pytest.mark.parametrize("name", ["128", "80", "110", "uint1", "uint2", "uint4"])
def test_tensor_core_keyswitch_bit_exact(name):
    """K2t (keyswitch_tc.cu): the key switch as a u8 x u8 -> s32 tcgen05 contraction over one-hot digits, four byte
    planes recombined mod 2^32 -- integer arithmetic, so the bar is equality with trgsw.identityKeySwitching
    (src/trgsw.zig:471-502) on every word, at row counts around the 128-row MMA tile and for every column-group width
    (n + 1 = 551, 631, 701: last groups of 40, 120 and 64 columns), and on the BASEBIT = 4 / 5 sets (UINT2 / UINT4: 2^BASEBIT one-hot
    bytes per digit, 6 / 3 digits per 96-byte K block)."""
    import tfhe_b200
    orc = O.Oracle(name); k = keys_for(name)
    c = tfhe_b200.Context(name, devices=[0])
    try:
        c.load_key(k.bsk, k.ksk, k.offset)
        rng = np.random.default_rng(5)
        c.set_tuning("ks_tc", 1)
        for B in (1, 127, 128, 129, 1000):
            lv1 = rng.integers(0, 2**32, (B, 1025), dtype=np.uint32)
            if B == 129:
                lv1[0] = 0; lv1[1] = 0xFFFFFFFF; lv1[2, :1024] = 0x55555555; lv1[3, :1024] = 0xAAAAAAAA   # digit patterns 0, 3, 1, 2
            got = c.keyswitch_batch(lv1)
            ref = orc.keyswitch_batch(lv1, k)
            assert (got == ref).all(), f"{name}: tensor-core key switch differs at B={B}: rows {np.nonzero((got != ref).any(axis=1))[0][:5]}"
        c.set_tuning("ks_tc", -1)
        assert (c.keyswitch_batch(lv1) == ref).all()          # scalar kernel, same bits
        c.set_tuning("ks_tc", 0)
        # whole gates through K1 -> K2t (automatic selection from 192 ciphertexts up); uint1's transform is exact only in exact mode
        if name.startswith("uint"):
            c.set_mode(tfhe_b200.MODE_EXACT)
        bits = rng.integers(0, 2, 300).astype(np.uint8)
        ca = orc.encrypt_bools(bits, k, 1); cb = orc.encrypt_bools(1 - bits, k, 2)
        out = c.gate_batch(O.NAND, ca, cb)
        sel = np.arange(0, 300, 13)
        assert (out[sel] == orc.gate_batch(O.NAND, ca[sel], cb[sel], k)).all()
    finally:
        c.close()


def test_a_failing_shard_reports_and_the_context_survives(ctx128, orc128, keys128):
    """error path of the host-batch driver: one device's shard fails (test hook) -> the call returns the error with its
    message, other shards are joined, and the next call on the same context gives the right bits again"""
    import tfhe_b200
    a, b, ca, cb = _enc_pairs(orc128, keys128, 64, seed=31)
    good = ctx128.gate_batch(O.NAND, ca, cb)
    ctx128.set_tuning("inject_fault", 1)
    with pytest.raises(tfhe_b200.TfheB200Error, match="injected fault"):
        ctx128.gate_batch(O.NAND, ca, cb)
    assert (ctx128.gate_batch(O.NAND, ca, cb) == good).all()
