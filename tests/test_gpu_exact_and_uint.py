"""Exact mode (reference DAG replay) and the large-digit UINT parameter sets (BASELINE config 4).

On UINT sets the FP64 external product is not an exact integer computation, so
coefficient-level parity needs TFHE_B200_MODE_EXACT; fast mode is checked at decode level there.
Known and documented: under the reference's u32-torus semantics UINT3+ (incl. UINT4) cannot decode f(m)
-- for those sets "parity" means GPU(exact) == oracle on identical keys, whatever they decode to."""
import numpy as np
import pytest

from conftest import keys_for
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _ctx(name, mode):
    import tfhe_b200
    k = keys_for(name)
    c = tfhe_b200.Context(name, devices=[0])
    c.load_key(k.bsk, k.ksk, k.offset)
    c.set_mode(mode)
    return c, O.Oracle(name), k


def test_exact_mode_128_equals_oracle_and_fast():
    import tfhe_b200
    c, orc, k = _ctx("128", tfhe_b200.MODE_EXACT)
    try:
        bits = np.array([0, 1, 1, 0, 1], np.uint8)
        ca = orc.encrypt_bools(bits, k, 1); cb = orc.encrypt_bools(bits[::-1].copy(), k, 2)
        lin = np.stack([orc.gate_linear(O.NAND, ca[i], cb[i]) for i in range(5)])
        ref = orc.blind_rotate_batch(lin, k)
        c.track_margin(True)
        got = c.blind_rotate_batch(lin)
        m = c.max_round_margin()
        c.track_margin(False)
        assert (got == ref).all()
        assert 0.05 < m < 0.25          # the reference DAG's own margin (~0.09), larger than the fast transform's
        c.set_mode(tfhe_b200.MODE_FAST)
        assert (c.blind_rotate_batch(lin) == ref).all()
        c.set_mode(tfhe_b200.MODE_EXACT)
        assert (c.gate_batch(O.NAND, ca, cb) == orc.gate_batch(O.NAND, ca, cb, k)).all()
    finally:
        c.close()


@pytest.mark.parametrize("name,modulus", [("uint1", 2), ("uint2", 4), ("uint4", 16)])
def test_exact_mode_uint_sets_bit_exact(name, modulus):
    import tfhe_b200
    c, orc, k = _ctx(name, tfhe_b200.MODE_EXACT)
    try:
        B = 6
        msgs = (np.arange(B) % modulus).astype(np.uint32)
        ct = orc.encrypt_lwe_messages(msgs, modulus, k, seed=5)
        table = np.array([(x * x + 1) % modulus for x in range(modulus)], np.uint32)
        tv = orc.lut_generate(table, modulus)
        ref_tr = orc.blind_rotate_batch(ct, k, tv)
        got_tr = c.blind_rotate_batch(ct, tv)
        assert (got_tr == ref_tr).all(), f"{(got_tr != ref_tr).sum()} coefficients differ"
        ref = orc.bootstrap_batch(ct, k, tv)
        got = c.bootstrap_batch(ct, tv)
        assert (got == ref).all()                      # includes the generic-base key switch (basebit 2/4/5)
        if name in ("uint1", "uint2"):                 # sets whose noise survives the u32 torus: f(m) must decode
            assert (orc.decrypt_lwe_messages(got, modulus, k) == table[msgs]).all()
    finally:
        c.close()


@pytest.mark.parametrize("name,modulus", [("uint1", 2), ("uint2", 4)])
def test_fast_mode_uint_sets_decode(name, modulus):
    """fast mode on large-digit sets: not coefficient-identical, but decodes the same f(m)"""
    import tfhe_b200
    c, orc, k = _ctx(name, tfhe_b200.MODE_FAST)
    try:
        B = 16
        msgs = (np.arange(B) % modulus).astype(np.uint32)
        ct = orc.encrypt_lwe_messages(msgs, modulus, k, seed=6)
        table = np.array([(x + 1) % modulus for x in range(modulus)], np.uint32)
        tv = orc.lut_generate(table, modulus)
        got = c.bootstrap_batch(ct, tv)
        assert (orc.decrypt_lwe_messages(got, modulus, k) == table[msgs]).all()
    finally:
        c.close()


@pytest.mark.parametrize("name,modulus", [("uint3", 8), ("uint5", 32), ("uint6", 64), ("uint7", 128), ("uint8", 256)])
def test_exact_mode_remaining_uint_sets_blind_rotation_bit_exact(name, modulus):
    """the UINT sets BASELINE config 4 does not quote (params.zig:180-375): n up to 1160, BGBIT 23 (UINT3) and
    22; real bootstrapping keys (their key-switching keys run to 1.8 GB and are covered by the structural test below), every CTA
    width of the exact kernel across a wave boundary, bit for bit against the oracle, where the reference's saturating
    float -> integer casts decide the result"""
    import tfhe_b200
    orc = O.Oracle(name); k = keys_for(name, with_ksk=False)
    c = tfhe_b200.Context(name, devices=[0])
    try:
        c.load_key(k.bsk, None, k.offset)
        c.set_mode(tfhe_b200.MODE_EXACT)
        B = 148 + 7
        rng = np.random.default_rng(12)
        msgs = rng.integers(0, modulus, B).astype(np.uint32)
        ct = orc.encrypt_lwe_messages(msgs, modulus, k, seed=9)
        tv = orc.lut_generate(np.array([(5 * x + 3) % modulus for x in range(modulus)], np.uint32), modulus)
        got = c.blind_rotate_batch(ct, tv)
        sel = np.array([0, 1, 5, 6, 147, 148, B - 1])
        ref = orc.blind_rotate_batch(ct[sel], k, tv)
        assert (got[sel] == ref).all(), f"{(got[sel] != ref).sum()} coefficients differ"
        for kct in (1, 4, 6):
            c.set_tuning("exact_kct", kct)
            assert (c.blind_rotate_batch(ct[:13], tv) == got[:13]).all(), kct
        c.set_tuning("exact_kct", 0)
        c.set_mode(tfhe_b200.MODE_FAST)      # fast mode: same sizes through the production kernel (not the parity mode on these sets)
        assert c.blind_rotate_batch(ct[:13], tv).shape == (13, 2, 1024)
    finally:
        c.close()


def test_keyswitch_generic_base_bit_exact():
    import tfhe_b200
    c, orc, k = _ctx("uint4", tfhe_b200.MODE_FAST)
    try:
        rng = np.random.default_rng(11)
        lv1 = rng.integers(0, 2**32, (19, 1025), dtype=np.uint32)
        assert (c.keyswitch_batch(lv1) == orc.keyswitch_batch(lv1, k)).all()
    finally:
        c.close()


def test_largest_lwe_dimension_set_fast_mode_runs():
    """UINT7/8 have the largest n (1160) and the widest key-switch base (2^7); keys are synthetic (random spectra),
    the check is structural: kernels accept the sizes and K2 matches the oracle on the generic-base path."""
    import tfhe_b200
    orc = O.Oracle("uint5")
    p = tfhe_b200.PARAM_SETS["uint5"]         # n = 1071, basebit 6: 196,608 KSK rows x 1072 u32 = 843 MB
    rng = np.random.default_rng(1)
    ksk = rng.integers(0, 2**32, (1024 * p.iks_t * (1 << p.basebit), p.n + 1), dtype=np.uint32)
    bsk = rng.standard_normal((p.n, 2 * p.L, 2, 1024)) * 1e6
    c = tfhe_b200.Context(p, devices=[0])
    try:
        c.load_key(bsk, ksk, 0x80000000)
        lv1 = rng.integers(0, 2**32, (9, 1025), dtype=np.uint32)
        keys = O.Keys(None, None, bsk, ksk, 0x80000000, None)
        assert (c.keyswitch_batch(lv1) == orc.keyswitch_batch(lv1, keys)).all()
        out = c.blind_rotate_batch(rng.integers(0, 2**32, (5, p.n + 1), dtype=np.uint32))
        assert out.shape == (5, 2, 1024)
    finally:
        c.close()


@pytest.mark.parametrize("name,modulus", [("uint4", 16), ("uint1", 2)])
def test_exact_register_blocked_equals_legacy_and_oracle(name, modulus):
    """the register-blocked exact kernel (default) at every CTA width, across a wave boundary, against the round-1
    one-CTA-per-ciphertext kernel and the oracle"""
    import tfhe_b200
    c, orc, k = _ctx(name, tfhe_b200.MODE_EXACT)
    try:
        B = 148 * 4 + 9
        rng = np.random.default_rng(2)
        msgs = rng.integers(0, modulus, B).astype(np.uint32)
        ct = orc.encrypt_lwe_messages(msgs, modulus, k, seed=8)
        tv = orc.lut_generate(np.array([(3 * x + 1) % modulus for x in range(modulus)], np.uint32), modulus)
        got = c.blind_rotate_batch(ct, tv)
        sel = np.array([0, 1, 2, 3, 4, 147, 148, 591, 592, 593, B - 1])
        assert (got[sel] == orc.blind_rotate_batch(ct[sel], k, tv)).all()
        c.set_tuning("exact_legacy", 1)
        legacy = c.blind_rotate_batch(ct, tv)
        c.set_tuning("exact_legacy", 0)
        assert (got == legacy).all()
        for kct in (1, 2, 3, 4, 6):      # 6: tensor-memory twiddles, X1 over X2, last-arriver ring
            c.set_tuning("exact_kct", kct)
            assert (c.blind_rotate_batch(ct[:23], tv) == got[:23]).all(), kct
        c.set_tuning("exact_kct", 4)
        assert (c.blind_rotate_batch(ct, tv) == got).all()      # the 601-item batch on the four-per-CTA kernel (default here: < 888)
        c.set_tuning("exact_kct", 0)
        c.track_margin(True)
        assert (c.blind_rotate_batch(ct[:9], tv) == got[:9]).all()
        assert c.max_round_margin() > 0.0
        c.track_margin(False)
        # per-item test vectors + extraction + key switch through the same kernel
        tvs = np.stack([orc.lut_generate(np.array([(x + s) % modulus for x in range(modulus)], np.uint32), modulus) for s in range(12)])
        assert (c.bootstrap_batch(ct[:12], tvs, tv_per_item=True) == orc.bootstrap_batch(ct[:12], k, tvs, tv_per_item=True)).all()
    finally:
        c.close()


@pytest.mark.parametrize("name,modulus", [("uint1", 2), ("uint2", 4), ("uint3", 8), ("uint4", 16), ("uint8", 256)])
def test_fast_mode_uint_instantiations_agree(name, modulus):
    """fast mode on the UINT sets is not the parity mode, but its kernels must agree with each other: the
    gadget-shape instantiations of the six-per-CTA kernel (L = 2 / BGBIT = 10, L = 1 / BGBIT = 18, 23, 22; full waves) against the
    generic four-per-CTA kernel and the round-1 kernels, word for word"""
    import tfhe_b200
    orc = O.Oracle(name); k = keys_for(name, with_ksk=False)
    c = tfhe_b200.Context(name, devices=[0])
    try:
        c.load_key(k.bsk, None, k.offset)
        B = 148 * 6 + 40
        rng = np.random.default_rng(12)
        ct = orc.encrypt_lwe_messages(rng.integers(0, modulus, B).astype(np.uint32), modulus, k, seed=3)
        tv = orc.lut_generate(np.arange(modulus, dtype=np.uint32), modulus)
        a = c.blind_rotate_batch(ct, tv)                 # 888 on the instantiation + a 40-ciphertext tail
        c.set_tuning("kct", 4)
        b = c.blind_rotate_batch(ct, tv)
        c.set_tuning("kct", 0); c.set_tuning("twt", -1)
        d = c.blind_rotate_batch(ct, tv)                 # round-1 kernels
        c.set_tuning("twt", 0)
        assert (a == b).all() and (a == d).all()
    finally:
        c.close()
