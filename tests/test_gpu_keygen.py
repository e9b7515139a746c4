"""Cloud-key generation on the device (tfhe_b200_keygen: key.genKeySwitchingKey / genBootstrappingKey, key.zig:148-212).
The reference's keys are clock-seeded and unpinned; what pins a generated key is what the reference's own key tests
check (key.zig:214-330: sizes, and that gates evaluated under it decrypt correctly) plus the defining relations of
every row, verified here with the CPU oracle on the exported reference-layout arrays."""
import numpy as np
import pytest

from conftest import keys_for
from oracle import oracle as O

pytestmark = pytest.mark.gpu
N = 1024
A_LV0, A_LV1 = 2.0e-5, 2.0e-8          # params.zig:350-375 (SECURITY_128_BIT): KSK_ALPHA, BSK_ALPHA


def _signed(x):
    return x.astype(np.uint32).view(np.int32).astype(np.float64)


def test_generated_key_rows_satisfy_their_definitions_and_gates_match_the_oracle():
    import tfhe_b200
    orc = O.Oracle("128"); ref = keys_for("128")
    s0, s1 = ref.s0, ref.s1
    c = tfhe_b200.Context("128", devices=[0])
    try:
        ck = c.keygen(s0, s1, seed=2024, ksk_alpha=A_LV0, bsk_alpha=A_LV1)
        n, L, t, base = orc.n, orc.L, 9, 4
        assert ck.bootstrapping_key.shape == (n, 2 * L, 2, N) and ck.key_switching_key.shape == (N * t * base, n + 1)
        assert ck.decomposition_offset == ref.offset == 0x82080000                      # key.zig:121-131

        # --- key switching key (key.zig:148-172): row (i, j, k) decrypts to k * s1[i] / 2^(2 (j + 1)) + N(0, alpha)
        ksk = ck.key_switching_key.reshape(N, t, base, n + 1)
        assert not ksk[:, :, 0].any()                                                   # k = 0 rows: never read, exported as zeros
        phase = (ksk[..., n].astype(np.int64) - (ksk[..., :n].astype(np.int64) * s0.astype(np.int64)).sum(-1)) & 0xFFFFFFFF
        i = np.arange(N)[:, None, None]; j = np.arange(t)[None, :, None]; k = np.arange(base)[None, None, :]
        want = ((k * s1.astype(np.int64)[i] * (1 << 32)) >> (2 * (j + 1))) & 0xFFFFFFFF
        err = _signed(((phase - want) & 0xFFFFFFFF)[:, :, 1:])
        sigma = A_LV0 * 2.0**32
        assert abs(err.std() / sigma - 1.0) < 0.03 and abs(err.mean()) < 0.05 * sigma and np.abs(err).max() < 6.5 * sigma
        masks = ksk[:, :, 1:, :n]
        assert abs(masks.astype(np.float64).mean() / 2.0**31 - 1.0) < 1e-3              # uniform u32 masks
        assert len(np.unique(masks[0, 0, 0])) == n

        # --- bootstrapping key (key.zig:182-212): each spectrum pair inverts to a TRLWE of 0 (+ gadget) under s1
        idx = [(0, 0), (1, 2), (5, 3), (n - 1, 2 * L - 1), (7, 1), (300, 4)]
        errs = []
        for (ii, r) in idx:
            a = O.fft1024(ck.bootstrapping_key[ii, r, 0]); b = O.fft1024(ck.bootstrapping_key[ii, r, 1])
            gadget = (int(s0[ii]) << (32 - 6 * ((r % L) + 1))) & 0xFFFFFFFF            # trgsw.zig:43-68
            a_plain = a.copy(); b_plain = b.copy()
            if r < L: a_plain[0] = (int(a[0]) - gadget) & 0xFFFFFFFF
            else: b_plain[0] = (int(b[0]) - gadget) & 0xFFFFFFFF
            prod = O.poly_mul_naive(a_plain, s1)                                          # trlwe.zig:54-61: b = noise + a (*) s1
            e = _signed((b_plain.astype(np.int64) - prod.astype(np.int64)) & 0xFFFFFFFF)
            errs.append(e)
            assert len(np.unique(a)) > N - 4
        errs = np.concatenate(errs)
        sigma1 = A_LV1 * 2.0**32
        assert abs(errs.std() / sigma1 - 1.0) < 0.08 and abs(errs.mean()) < 0.1 * sigma1, (errs.std(), sigma1)

        # --- the key works, and the device layouts the kernels read are the exported key: GPU == oracle on it, bit for bit
        keys = O.Keys(s0, s1, ck.bootstrapping_key, ck.key_switching_key, ck.decomposition_offset, ref.testvec)
        rng = np.random.default_rng(8)
        B = 24
        abits = rng.integers(0, 2, B).astype(np.uint8); bbits = rng.integers(0, 2, B).astype(np.uint8)
        ca = orc.encrypt_bools(abits, keys, 31); cb = orc.encrypt_bools(bbits, keys, 32)
        out = c.gate_batch(O.NAND, ca, cb)
        assert (orc.decrypt_bools(out, keys) == 1 - (abits & bbits)).all()
        assert (out[:4] == orc.gate_batch(O.NAND, ca[:4], cb[:4], keys)).all()
        c.set_mode(tfhe_b200.MODE_EXACT)                                                # exact mode reads the reference-layout copy
        assert (c.gate_batch(O.XOR, ca[:3], cb[:3]) == orc.gate_batch(O.XOR, ca[:3], cb[:3], keys)).all()
        c.set_mode(tfhe_b200.MODE_FAST)

        # --- reproducible from the seed; a different seed gives a different key
        ck2 = c.keygen(s0, s1, seed=2024, ksk_alpha=A_LV0, bsk_alpha=A_LV1)
        assert (ck2.bootstrapping_key == ck.bootstrapping_key).all() and (ck2.key_switching_key == ck.key_switching_key).all()
        ck3 = c.keygen(s0, s1, seed=2025, ksk_alpha=A_LV0, bsk_alpha=A_LV1)
        assert (ck3.key_switching_key[1] != ck.key_switching_key[1]).any()
        out3 = c.gate_batch(O.NAND, ca, cb)                                             # same secret key: still decrypts
        assert (orc.decrypt_bools(out3, keys) == 1 - (abits & bbits)).all()
    finally:
        c.close()
