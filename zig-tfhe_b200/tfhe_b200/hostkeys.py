"""Host-side key generation, encryption and decryption (numpy).

In the reference these stay on the host (key.zig:41-212, tlwe.zig:34-117, trlwe.zig:30-64,
trgsw.zig:35-91, utils.zig:28-116) and are NOT part of the accelerated path; this module mirrors
them so that bench.py and the examples can make seeded synthetic keys / ciphertexts and check decrypted
outputs without touching the test oracle.  Vectorised restatement, same distributions and layouts;
it does not promise the oracle's bit pattern (different PRNG stream).

Spectra are produced in the reference's own format (fft.zig:293-366: twist, 512-point forward DFT,
x2, split re|im) with numpy's FFT, which is what tfhe_b200_load_key expects.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import CloudKey, PARAM_SETS, Params

N = 1024
# params.zig:70-375 noise levels (alpha_lv0 = KSK/TLWE noise, alpha_lv1 = BSK/TRLWE noise)
ALPHAS = {
    "80": (5.0e-5, 3.73e-8), "110": (3.0517578125e-05, 2.9802322387695313e-8), "128": (2.0e-5, 2.0e-8),
    "uint1": (2.0e-05, 2.0e-08), "uint2": (0.00002120846893069971872305794214, 0.00000000000231841227527049948463),
    "uint3": (0.00000251676160959795544987084234, 0.00000000000000022204460492503131),
    "uint4": (0.00000251676160959795544987084234, 0.00000000000000022204460492503131),
    "uint5": (7.088226765410429399593757e-08, 2.2204460492503131e-17), "uint6": (7.088226765410429399593757e-08, 2.2204460492503131e-17),
    "uint7": (1.966220007498402695211596e-08, 2.2204460492503131e-17), "uint8": (1.966220007498402695211596e-08, 2.2204460492503131e-17),
}

_TWIST = np.exp(1j * np.pi * np.arange(N // 2) / N)


def f64_to_torus(d):
    """utils.zig:28-33 (vectorised): mod 1, scale by 2^32, clamp, truncate."""
    d = np.asarray(d, dtype=np.float64)
    a = np.fmod(d, 1.0)
    normalized = np.where(d < 0.0, np.fmod(a + 1.0, 1.0), a)
    t = np.clip(normalized * 4294967296.0, 0.0, 4294967295.0)
    return t.astype(np.uint64).astype(np.uint32)


def spectrum(poly_u32):
    """reference forward transform ifft1024 (fft.zig:293-366) of [..., 1024] u32 -> [..., 1024] f64."""
    p = np.asarray(poly_u32, dtype=np.uint32).view(np.int32).astype(np.float64)
    z = (p[..., : N // 2] + 1j * p[..., N // 2:]) * _TWIST
    Z = np.fft.fft(z, axis=-1)
    return np.concatenate([2.0 * Z.real, 2.0 * Z.imag], axis=-1)


def negacyclic_mul_binary(a_u32, s_bits):
    """a(X) * s(X) mod (X^N+1, 2^32) for binary s, exact (FFT in f64 on 16-bit limbs)."""
    a = np.asarray(a_u32, dtype=np.uint32)
    s = np.asarray(s_bits, dtype=np.float64)
    tw = np.exp(1j * np.pi * np.arange(N) / N)
    S = np.fft.fft(s * tw)
    out = np.zeros(a.shape, dtype=np.uint64)
    for shift in (0, 16):
        limb = ((a >> shift) & 0xFFFF).astype(np.float64)
        prod = np.fft.ifft(np.fft.fft(limb * tw, axis=-1) * S, axis=-1) * np.conj(tw)
        out += (np.rint(prod.real).astype(np.int64).astype(np.uint64) << np.uint64(shift))
    return (out & np.uint64(0xFFFFFFFF)).astype(np.uint32)


@dataclass
class SecretKey:
    """key.SecretKey (key.zig:34-58): binary lv0[n] and lv1[N]."""
    key_lv0: np.ndarray
    key_lv1: np.ndarray


def gen_secret_key(params: Params, rng) -> SecretKey:
    return SecretKey(rng.integers(0, 2, params.n, dtype=np.uint32), rng.integers(0, 2, N, dtype=np.uint32))


def tlwe_encrypt_f64(mu, alpha, key, rng):
    """tlwe.zig:34-50 over a batch: mu [B] -> [B][len(key)+1]."""
    mu = np.atleast_1d(np.asarray(mu, dtype=np.float64))
    B, n = mu.shape[0], len(key)
    out = np.empty((B, n + 1), np.uint32)
    out[:, :n] = rng.integers(0, 2**32, (B, n), dtype=np.uint32)
    inner = (out[:, :n] * key[None, :]).sum(axis=1, dtype=np.uint32)
    noise = f64_to_torus(rng.normal(0.0, alpha, B)) + f64_to_torus(mu)
    out[:, n] = inner + noise
    return out


def encrypt_bools(bits, params: Params, sk: SecretKey, rng):
    """tlwe.zig:53-56"""
    a0 = ALPHAS[params.name][0]
    return tlwe_encrypt_f64(np.where(np.asarray(bits) != 0, 0.125, -0.125), a0, sk.key_lv0, rng)


def decrypt_bools(ct, sk: SecretKey, level=0):
    """tlwe.zig:59-69"""
    key = sk.key_lv0 if level == 0 else sk.key_lv1
    ct = np.asarray(ct, dtype=np.uint32).reshape(-1, len(key) + 1)
    phase = ct[:, -1] - (ct[:, :-1] * key[None, :]).sum(axis=1, dtype=np.uint32)
    return (phase.view(np.int32) >= 0).astype(np.uint8)


def gen_decomposition_offset(params: Params) -> int:
    """key.zig:121-131"""
    off = 0
    for i in range(params.L):
        off += (1 << (params.bgbit - 1)) << (32 - (i + 1) * params.bgbit)
    return off & 0xFFFFFFFF


def gen_key_switching_key(params: Params, sk: SecretKey, rng):
    """key.zig:148-172: [N*t*base][n+1]; k = 0 rows left zero (never read)."""
    base, t, n = 1 << params.basebit, params.iks_t, params.n
    a0 = ALPHAS[params.name][0]
    ksk = np.zeros((N, t, base, n + 1), np.uint32)
    for j in range(t):
        for k in range(1, base):
            mu = (k * sk.key_lv1.astype(np.float64)) / float(1 << ((j + 1) * params.basebit))
            ksk[:, j, k, :] = tlwe_encrypt_f64(mu, a0, sk.key_lv0, rng)
    return ksk.reshape(N * t * base, n + 1)


def gen_reencryption_key(params: Params, key_from, key_to, rng, basebit=None, t=None):
    """proxy_reenc.ProxyReencryptionKey.newSymmetric (proxy_reenc.zig:205-262): entry (i, j, k) encrypts
    k * key_from[i] / 2^((j+1) basebit) under key_to; [n*t*base][n+1], k = 0 rows left zero (never read)."""
    basebit = params.basebit if basebit is None else basebit
    t = params.iks_t if t is None else t
    base, n = 1 << basebit, params.n
    a0 = ALPHAS[params.name][0]
    key = np.zeros((n, t, base, n + 1), np.uint32)
    for j in range(t):
        for k in range(1, base):
            mu = (k * np.asarray(key_from, np.float64)) / float(1 << ((j + 1) * basebit))
            key[:, j, k, :] = tlwe_encrypt_f64(mu, a0, np.asarray(key_to, np.uint32), rng)
    return key.reshape(n * t * base, n + 1)


def gen_bootstrapping_key(params: Params, sk: SecretKey, rng):
    """key.zig:182-212 (TRGSW(s0_i) in FFT form): f64 [n][2L][2][N]."""
    n, L = params.n, params.L
    a1 = ALPHAS[params.name][1]
    bsk = np.empty((n, 2 * L, 2, N), np.float64)
    gadget = [int(f64_to_torus(float(1 << params.bgbit) ** -(i + 1))) for i in range(L)]   # trgsw.zig:47-51
    rows_per_block = 64
    for i0 in range(0, n, rows_per_block):
        i1 = min(n, i0 + rows_per_block)
        m = i1 - i0
        a = rng.integers(0, 2**32, (m, 2 * L, N), dtype=np.uint32)          # trlwe.zig:40-42
        b = f64_to_torus(rng.normal(0.0, a1, (m, 2 * L, N)))                 # trlwe.zig:45-52 (mu = 0)
        b = b + negacyclic_mul_binary(a, sk.key_lv1)                         # trlwe.zig:55-61
        for r in range(L):                                                   # trgsw.zig:65-68
            g = (sk.key_lv0[i0:i1] * np.uint32(gadget[r])).astype(np.uint32)
            a[:, r, 0] += g
            b[:, r + L, 0] += g
        bsk[i0:i1, :, 0, :] = spectrum(a)
        bsk[i0:i1, :, 1, :] = spectrum(b)
    return bsk


def gen_cloud_key(params: Params | str, seed: int = 1, with_ksk: bool = True):
    """key.CloudKey.new (key.zig:70-77) from a fresh seeded secret key."""
    params = PARAM_SETS[params] if isinstance(params, str) else params
    rng = np.random.default_rng(seed)
    sk = gen_secret_key(params, rng)
    bsk = gen_bootstrapping_key(params, sk, rng)
    ksk = gen_key_switching_key(params, sk, rng) if with_ksk else None
    tv = np.zeros((2, N), np.uint32)
    tv[1, :] = 0x20000000   # key.zig:134-145
    return sk, CloudKey(bsk, ksk, gen_decomposition_offset(params), tv)
