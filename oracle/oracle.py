"""ctypes loader for the CPU oracle (oracle/libtfhe_oracle.so).

TEST INFRASTRUCTURE ONLY.  Allowed importers: tests/, __graft_entry__.smoke(), and
bench.py's cpu_baseline / --impl reference legs.  The product package
(zig-tfhe_b200/tfhe_b200) must never import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtfhe_oracle.so")

NAND, OR, AND, XOR, XNOR, NOR, ANDNY, ANDYN, ORNY, ORYN = range(10)
N = 1024


class OrcParams(C.Structure):
    _fields_ = [("n", C.c_int32), ("N", C.c_int32), ("L", C.c_int32), ("bgbit", C.c_int32),
                ("basebit", C.c_int32), ("iks_t", C.c_int32), ("alpha_lv0", C.c_double), ("alpha_lv1", C.c_double)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "tfhe_oracle.cpp")
    hdr = os.path.join(_HERE, "tfhe_oracle.h")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libtfhe_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_bsk_len.restype = C.c_size_t
        _lib.orc_ksk_len.restype = C.c_size_t
        _lib.orc_f64_to_torus.restype = C.c_uint32
        _lib.orc_f64_to_torus.argtypes = [C.c_double]
        _lib.orc_torus_to_f64.restype = C.c_double
        _lib.orc_torus_to_f64.argtypes = [C.c_uint32]
        _lib.orc_decomposition_offset.restype = C.c_uint32
        _lib.orc_tlwe_phase.restype = C.c_uint32
        _lib.orc_lut_encode.restype = C.c_uint32
        _lib.orc_lut_encode.argtypes = [C.c_uint32, C.c_uint32]
        _lib.orc_lut_decode.restype = C.c_uint32
        _lib.orc_lut_decode.argtypes = [C.c_uint32, C.c_uint32]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def hardware_threads() -> int:
    return int(lib().orc_hardware_threads())


def f64_to_torus(d: float) -> int:
    return int(lib().orc_f64_to_torus(float(d)))


def ifft1024(x):
    x = _u32(x); out = np.empty(N, np.float64)
    lib().orc_ifft1024(_p(x), _p(out)); return out


def fft1024(s, with_margin=False):
    s = _f64(s); out = np.empty(N, np.uint32)
    if with_margin:
        m = C.c_double(0.0)
        lib().orc_fft1024_margin(_p(s), _p(out), C.byref(m)); return out, m.value
    lib().orc_fft1024(_p(s), _p(out)); return out


def poly_mul_fft(a, b):
    a = _u32(a); b = _u32(b); out = np.empty(N, np.uint32)
    lib().orc_poly_mul_fft(_p(a), _p(b), _p(out)); return out


def poly_mul_naive(a, b):
    a = _u32(a); b = _u32(b); out = np.empty(len(a), np.uint32)
    lib().orc_poly_mul_naive(_p(a), _p(b), _p(out), C.c_int(len(a))); return out


def radix2_fft(z, inverse=False):
    z = np.asarray(z, dtype=np.complex128)
    d = np.ascontiguousarray(np.stack([z.real, z.imag], axis=1).reshape(-1))
    lib().orc_radix2_fft(_p(d), C.c_int(len(z)), C.c_int(1 if inverse else 0))
    d = d.reshape(-1, 2)
    return d[:, 0] + 1j * d[:, 1]


def poly_mul_with_xk(a, k):
    a = _u32(a); out = np.empty(N, np.uint32)
    lib().orc_poly_mul_with_xk(_p(a), C.c_size_t(k), _p(out)); return out


def lut_encode(m, modulus):
    return int(lib().orc_lut_encode(int(m), int(modulus)))


def lut_decode(v, modulus):
    return int(lib().orc_lut_decode(int(v), int(modulus)))


@dataclass
class Keys:
    s0: np.ndarray
    s1: np.ndarray
    bsk: np.ndarray          # [n][2L][2][N] f64
    ksk: np.ndarray | None   # [N*t*base][n+1] u32
    offset: int
    testvec: np.ndarray      # [2][N] u32


class Oracle:
    """One parameter set + the reference semantics bound to it."""

    def __init__(self, name: str = "128"):
        self.name = name
        self.p = OrcParams()
        if lib().orc_get_params(name.encode(), C.byref(self.p)) != 0:
            raise ValueError(f"unknown parameter set {name!r}")
        self.n, self.L, self.bgbit = self.p.n, self.p.L, self.p.bgbit
        self.basebit, self.iks_t = self.p.basebit, self.p.iks_t
        self.N = N
        self._pp = C.byref(self.p)

    # -- sizes
    @property
    def bsk_len(self):
        return int(lib().orc_bsk_len(self._pp))

    @property
    def ksk_len(self):
        return int(lib().orc_ksk_len(self._pp))

    @property
    def ksk_rows(self):
        return N * self.iks_t * (1 << self.basebit)

    # -- keys
    def keygen(self, seed: int = 1, with_ksk: bool = True, noiseless_bsk: bool = False) -> Keys:
        s0 = np.empty(self.n, np.uint32); s1 = np.empty(N, np.uint32)
        lib().orc_gen_secret(self._pp, C.c_uint64(seed), _p(s0), _p(s1))
        bsk = np.empty(self.bsk_len, np.float64)
        fn = lib().orc_gen_bsk_noiseless if noiseless_bsk else lib().orc_gen_bsk
        fn(self._pp, C.c_uint64(seed), _p(s0), _p(s1), _p(bsk))
        ksk = None
        if with_ksk:
            ksk = np.empty(self.ksk_len, np.uint32)
            lib().orc_gen_ksk(self._pp, C.c_uint64(seed), _p(s0), _p(s1), _p(ksk))
            ksk = ksk.reshape(self.ksk_rows, self.n + 1)
        tv = np.empty(2 * N, np.uint32)
        lib().orc_gen_testvec(self._pp, _p(tv))
        off = int(lib().orc_decomposition_offset(self._pp))
        return Keys(s0, s1, bsk.reshape(self.n, 2 * self.L, 2, N), ksk, off, tv.reshape(2, N))

    # -- encrypt / decrypt
    def encrypt_bools(self, bits, keys: Keys, seed: int = 42):
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        out = np.empty((len(bits), self.n + 1), np.uint32)
        lib().orc_tlwe_encrypt_bools(self._pp, C.c_uint64(seed), _p(bits), C.c_size_t(len(bits)), _p(keys.s0), _p(out))
        return out

    def decrypt_bools(self, ct, keys: Keys, level: int = 0):
        ct = _u32(ct); key = keys.s0 if level == 0 else keys.s1
        ct2 = ct.reshape(-1, len(key) + 1)
        bits = np.empty(ct2.shape[0], np.uint8)
        lib().orc_tlwe_decrypt_bools(self._pp, _p(ct2), C.c_size_t(ct2.shape[0]), _p(key), C.c_int(len(key)), _p(bits))
        return bits

    def phase(self, ct, keys: Keys, level: int = 0):
        ct = _u32(ct); key = keys.s0 if level == 0 else keys.s1
        ct2 = ct.reshape(-1, len(key) + 1)
        return np.array([lib().orc_tlwe_phase(_p(ct2[i]), _p(key), C.c_int(len(key))) for i in range(ct2.shape[0])], np.uint32)

    def encrypt_lwe_messages(self, msgs, modulus, keys: Keys, seed: int = 42):
        msgs = _u32(msgs)
        out = np.empty((len(msgs), self.n + 1), np.uint32)
        lib().orc_tlwe_encrypt_lwe_messages(self._pp, C.c_uint64(seed), _p(msgs), C.c_size_t(len(msgs)), C.c_uint32(modulus), _p(keys.s0), _p(out))
        return out

    def decrypt_lwe_messages(self, ct, modulus, keys: Keys, level: int = 0):
        ct = _u32(ct); key = keys.s0 if level == 0 else keys.s1
        ct2 = ct.reshape(-1, len(key) + 1)
        out = np.empty(ct2.shape[0], np.uint32)
        lib().orc_tlwe_decrypt_lwe_messages(self._pp, _p(ct2), C.c_size_t(ct2.shape[0]), C.c_uint32(modulus), _p(key), C.c_int(len(key)), _p(out))
        return out

    def trlwe_encrypt_f64(self, mu, keys: Keys, seed: int = 7, alpha=None):
        mu = _f64(mu); out = np.empty((2, N), np.uint32)
        lib().orc_trlwe_encrypt_f64(self._pp, C.c_uint64(seed), _p(mu), C.c_double(self.p.alpha_lv1 if alpha is None else alpha), _p(keys.s1), _p(out))
        return out

    def trlwe_encrypt_bools(self, bits, keys: Keys, seed: int = 7):
        return self.trlwe_encrypt_f64(np.where(np.asarray(bits) != 0, 0.125, -0.125), keys, seed)

    def trlwe_phase(self, ct, keys: Keys):
        ct = _u32(ct); out = np.empty(N, np.uint32)
        lib().orc_trlwe_phase(self._pp, _p(ct), _p(keys.s1), _p(out)); return out

    def trlwe_decrypt_bools(self, ct, keys: Keys):
        return (self.trlwe_phase(ct, keys).view(np.int32) >= 0).astype(np.uint8)

    def trgsw_encrypt_fft(self, msg: int, keys: Keys, seed: int = 9):
        out = np.empty((2 * self.L, 2, N), np.float64)
        lib().orc_trgsw_encrypt_fft(self._pp, C.c_uint64(seed), C.c_uint32(msg), C.c_double(self.p.alpha_lv1), _p(keys.s1), _p(out))
        return out

    # -- hot path
    def gate_linear(self, op, a, b):
        a = _u32(a); b = _u32(b); out = np.empty(self.n + 1, np.uint32)
        lib().orc_gate_linear(self._pp, C.c_int(op), _p(a), _p(b), _p(out)); return out

    def gate_not(self, a):
        a = _u32(a); out = np.empty(self.n + 1, np.uint32)
        lib().orc_gate_not(self._pp, _p(a), _p(out)); return out

    def gate_constant(self, value: bool):
        out = np.empty(self.n + 1, np.uint32)
        lib().orc_gate_constant(self._pp, C.c_int(1 if value else 0), _p(out)); return out

    def decomposition(self, trlwe, offset):
        t = _u32(trlwe); out = np.empty((2 * self.L, N), np.uint32)
        lib().orc_decomposition(self._pp, _p(t), C.c_uint32(offset), _p(out)); return out

    def external_product(self, trgsw_fft, trlwe, offset, with_margin=False):
        g = _f64(trgsw_fft); t = _u32(trlwe); out = np.empty((2, N), np.uint32); m = C.c_double(0.0)
        lib().orc_external_product(self._pp, _p(g), _p(t), C.c_uint32(offset), _p(out), C.byref(m))
        return (out, m.value) if with_margin else out

    def external_product_int(self, trgsw_fft, trlwe, offset):
        g = _f64(trgsw_fft); t = _u32(trlwe); out = np.empty((2, N), np.uint32)
        lib().orc_external_product_int(self._pp, _p(g), _p(t), C.c_uint32(offset), _p(out)); return out

    def cmux(self, in1, in2, cond_fft, offset):
        a = _u32(in1); b = _u32(in2); g = _f64(cond_fft); out = np.empty((2, N), np.uint32); m = C.c_double(0.0)
        lib().orc_cmux(self._pp, _p(a), _p(b), _p(g), C.c_uint32(offset), _p(out), C.byref(m)); return out

    def blind_rotate(self, src, keys: Keys, testvec=None, trace=False, with_margin=False):
        s = _u32(src); out = np.empty((2, N), np.uint32)
        tv = _u32(testvec) if testvec is not None else None
        tr = np.empty((self.n, 2, N), np.uint32) if trace else None
        m = C.c_double(0.0)
        lib().orc_blind_rotate(self._pp, _p(s), _p(keys.bsk), C.c_uint32(keys.offset), _p(tv), _p(out), _p(tr), C.byref(m))
        res = [out]
        if trace: res.append(tr)
        if with_margin: res.append(m.value)
        return res[0] if len(res) == 1 else tuple(res)

    def sample_extract_index(self, trlwe, k=0):
        t = _u32(trlwe); out = np.empty(N + 1, np.uint32)
        lib().orc_sample_extract_index(self._pp, _p(t), C.c_int(k), _p(out)); return out

    def sample_extract_index2(self, trlwe, k=0):
        t = _u32(trlwe); out = np.empty(self.n + 1, np.uint32)
        lib().orc_sample_extract_index2(self._pp, _p(t), C.c_int(k), _p(out)); return out

    def identity_key_switching(self, lv1, keys: Keys):
        v = _u32(lv1); out = np.empty(self.n + 1, np.uint32)
        lib().orc_identity_key_switching(self._pp, _p(v), _p(keys.ksk), _p(out)); return out

    def bootstrap(self, ct, keys: Keys, testvec=None):
        c = _u32(ct); out = np.empty(self.n + 1, np.uint32)
        tv = _u32(testvec) if testvec is not None else None
        lib().orc_bootstrap(self._pp, _p(c), _p(keys.bsk), _p(keys.ksk), C.c_uint32(keys.offset), _p(tv), _p(out)); return out

    def gate(self, op, a, b, keys: Keys):
        return self.bootstrap(self.gate_linear(op, a, b), keys)

    def gate_batch(self, op, a, b, keys: Keys, nthreads=None):
        a = _u32(a); b = _u32(b); B = a.shape[0]
        out = np.empty((B, self.n + 1), np.uint32)
        ops = None
        if not np.isscalar(op):
            ops = np.ascontiguousarray(op, dtype=np.int32); op = 0
        lib().orc_gate_batch(self._pp, C.c_int(int(op)), _p(ops), _p(a), _p(b), _p(out), C.c_size_t(B), _p(keys.bsk), _p(keys.ksk),
                             C.c_uint32(keys.offset), C.c_int(nthreads or hardware_threads()))
        return out

    def bootstrap_batch(self, ct, keys: Keys, testvec=None, tv_per_item=False, nthreads=None):
        c = _u32(ct); B = c.shape[0]; out = np.empty((B, self.n + 1), np.uint32)
        tv = _u32(testvec) if testvec is not None else None
        lib().orc_bootstrap_batch(self._pp, _p(c), _p(out), C.c_size_t(B), _p(keys.bsk), _p(keys.ksk), C.c_uint32(keys.offset), _p(tv),
                                  C.c_int(1 if tv_per_item else 0), C.c_int(nthreads or hardware_threads()))
        return out

    def blind_rotate_batch(self, ct, keys: Keys, testvec=None, tv_per_item=False, nthreads=None):
        c = _u32(ct); B = c.shape[0]; out = np.empty((B, 2, N), np.uint32)
        tv = _u32(testvec) if testvec is not None else None
        lib().orc_blind_rotate_batch(self._pp, _p(c), _p(out), C.c_size_t(B), _p(keys.bsk), C.c_uint32(keys.offset), _p(tv),
                                     C.c_int(1 if tv_per_item else 0), C.c_int(nthreads or hardware_threads()))
        return out

    def keyswitch_batch(self, lv1, keys: Keys, nthreads=None):
        v = _u32(lv1); B = v.shape[0]; out = np.empty((B, self.n + 1), np.uint32)
        lib().orc_keyswitch_batch(self._pp, _p(v), _p(out), C.c_size_t(B), _p(keys.ksk), C.c_int(nthreads or hardware_threads()))
        return out

    def gen_reenc_key(self, key_from, key_to, seed=3, basebit=None, t=None):
        basebit = self.basebit if basebit is None else basebit
        t = self.iks_t if t is None else t
        out = np.empty((self.n * t * (1 << basebit), self.n + 1), np.uint32)
        lib().orc_gen_reenc_key(self._pp, C.c_uint64(seed), _p(_u32(key_from)), _p(_u32(key_to)), C.c_int(basebit), C.c_int(t), _p(out))
        return out

    def reencrypt(self, ct, key, basebit=None, t=None):
        basebit = self.basebit if basebit is None else basebit
        t = self.iks_t if t is None else t
        ct = _u32(ct).reshape(-1, self.n + 1); key = _u32(key)
        out = np.empty_like(ct)
        for i in range(ct.shape[0]):
            lib().orc_reencrypt(self._pp, _p(ct[i]), _p(key), C.c_int(basebit), C.c_int(t), _p(out[i]))
        return out

    def lut_generate(self, table, modulus):
        t = _u32(table); out = np.empty((2, N), np.uint32)
        lib().orc_lut_generate(self._pp, _p(t), C.c_uint32(modulus), _p(out)); return out
