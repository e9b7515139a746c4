"""Host-side binding of libtfhe_b200 (include/tfhe_b200.h) for tests and bench.py.

The reference's host language is Zig (no toolchain in this image), so the drop-in shim lives in
zig-tfhe_b200/zig/ and the C++ mirror in zig-tfhe_b200/host/.  This module is the Python ctypes
view of the same C ABI, shaped after the reference's own API for this path:

  reference (Zig)                                  here
  -----------------------------------------------  ------------------------------------------
  params.SECURITY_128_BIT ... (params.zig:70-375)  PARAM_SETS["128"] ...
  key.CloudKey (key.zig:61-65)                     CloudKey
  gates.Gates{bootstrap} (gates.zig:25-151)        Gates(ctx)  .nand/.and_/.or_/.xor/... on batches
  gates.batchNand ... (gates.zig:244-295)          batch_nand ... (same names, snake case)
  VanillaBootstrap.bootstrap (vanilla.zig:38-52)   GpuBootstrap.bootstrap
  trgsw.batchBlindRotate (trgsw.zig:415-421)       Context.blind_rotate_batch

There is no CPU fallback: if the CUDA library is missing or no sm_100 device is present every
entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# TFHE_B200_LIB: another build of the same library (tools/build_diag.sh makes a diagnostic one); never a different implementation
LIB_PATH = os.environ.get("TFHE_B200_LIB") or os.path.join(os.path.dirname(_HERE), "libtfhe_b200.so")

NAND, OR, AND, XOR, XNOR, NOR, ANDNY, ANDYN, ORNY, ORYN = range(10)
GATE_NAMES = ["nand", "or", "and", "xor", "xnor", "nor", "andny", "andyn", "orny", "oryn"]
MODE_FAST, MODE_EXACT = 0, 1
N = 1024

STATUS = {0: "ok", 1: "invalid argument", 2: "no sm_100 CUDA device", 3: "CUDA error", 4: "no key loaded", 5: "not implemented", 6: "I/O error"}


class TfheB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"tfhe_b200 error {code} ({STATUS.get(code, '?')}): {msg}")
        self.code = code


class NotImplementedOnDevice(TfheB200Error):
    """mirrors Zig's error.NotImplemented"""


class _Params(C.Structure):
    _fields_ = [("n", C.c_int32), ("N", C.c_int32), ("L", C.c_int32), ("bgbit", C.c_int32), ("basebit", C.c_int32), ("iks_t", C.c_int32)]


@dataclass(frozen=True)
class Params:
    """runtime mirror of params.SecurityParams (params.zig:36-67)"""
    name: str
    n: int
    L: int
    bgbit: int
    basebit: int
    iks_t: int
    N: int = 1024


# params.zig:70-375
PARAM_SETS = {
    "80": Params("80", 550, 3, 6, 2, 7),
    "110": Params("110", 630, 3, 6, 2, 8),
    "128": Params("128", 700, 3, 6, 2, 9),
    "uint1": Params("uint1", 700, 2, 10, 2, 8),
    "uint2": Params("uint2", 687, 1, 18, 4, 3),
    "uint3": Params("uint3", 820, 1, 23, 6, 2),
    "uint4": Params("uint4", 820, 1, 22, 5, 3),
    "uint5": Params("uint5", 1071, 1, 22, 6, 3),
    "uint6": Params("uint6", 1071, 1, 22, 6, 3),
    "uint7": Params("uint7", 1160, 1, 22, 7, 3),
    "uint8": Params("uint8", 1160, 1, 22, 7, 3),
}

_lib = None


def load_library():
    """dlopen the CUDA library; fails loudly if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} not built: run `python __graft_entry__.py` (nvcc, sm_100a)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, u32, sz = C.c_void_p, C.c_int, C.c_uint32, C.c_size_t
    sig = {
        "tfhe_b200_create": (i32, [vp, vp, i32, vp]),
        "tfhe_b200_destroy": (None, [vp]),
        "tfhe_b200_last_error": (C.c_char_p, [vp]),
        "tfhe_b200_num_devices": (i32, [vp]),
        "tfhe_b200_version": (C.c_char_p, []),
        "tfhe_b200_load_key": (i32, [vp, vp, vp, sz, u32]),
        "tfhe_b200_load_key_device": (i32, [vp, i32, vp, vp, u32]),
        "tfhe_b200_set_mode": (i32, [vp, i32]),
        "tfhe_b200_gate_batch": (i32, [vp, i32, vp, vp, vp, sz]),
        "tfhe_b200_gate_batch_ops": (i32, [vp, vp, vp, vp, vp, sz]),
        "tfhe_b200_bootstrap_batch": (i32, [vp, vp, vp, sz, vp, i32]),
        "tfhe_b200_bootstrap_no_keyswitch_batch": (i32, [vp, vp, vp, sz]),
        "tfhe_b200_lut_bootstrap_many_batch": (i32, [vp, vp, vp, sz, vp, i32, i32]),
        "tfhe_b200_blind_rotate_batch": (i32, [vp, vp, vp, sz, vp, i32]),
        "tfhe_b200_keyswitch_batch": (i32, [vp, vp, vp, sz]),
        "tfhe_b200_blind_rotate_extract_batch": (i32, [vp, vp, vp, sz]),
        "tfhe_b200_not_batch": (i32, [vp, vp, vp, sz]),
        "tfhe_b200_load_reencryption_key": (i32, [vp, vp, i32, i32]),
        "tfhe_b200_reencrypt_batch": (i32, [vp, vp, vp, sz]),
        "tfhe_b200_gate_batch_device": (i32, [vp, i32, i32, vp, vp, vp, vp, sz]),
        "tfhe_b200_bootstrap_batch_device": (i32, [vp, i32, vp, vp, sz, vp, i32]),
        "tfhe_b200_blind_rotate_batch_device": (i32, [vp, i32, vp, vp, sz, vp, i32]),
        "tfhe_b200_keyswitch_batch_device": (i32, [vp, i32, vp, vp, sz]),
        "tfhe_b200_stream": (vp, [vp, i32]),
        "tfhe_b200_sync": (i32, [vp]),
        "tfhe_b200_track_margin": (i32, [vp, i32]),
        "tfhe_b200_max_round_margin": (C.c_double, [vp, i32]),
        "tfhe_b200_launch_count": (C.c_uint64, [vp]),
        "tfhe_b200_set_tuning": (i32, [vp, C.c_char_p, i32]),
        "tfhe_b200_measure_fp64_tflops": (C.c_double, [vp, i32]),
        "tfhe_b200_last_kernel_ms": (C.c_double, [vp, i32, i32]),
        "tfhe_b200_keygen": (i32, [vp, vp, vp, C.c_uint64, C.c_double, C.c_double, vp, vp]),
        "tfhe_b200_decomposition_offset": (u32, [vp]),
        "tfhe_b200_lut_bootstrap_batch": (i32, [vp, vp, vp, sz, vp, i32, i32]),
        "tfhe_b200_lut_generate": (i32, [vp, vp, i32, vp]),
        "tfhe_b200_circuit_create": (i32, [vp, vp, sz, sz, vp, sz, vp]),
        "tfhe_b200_circuit_destroy": (None, [vp]),
        "tfhe_b200_circuit_plan": (i32, [vp, sz, sz, vp, sz, vp, vp, vp]),
        "tfhe_b200_circuit_info": (i32, [vp, vp, vp, vp]),
        "tfhe_b200_circuit_run": (i32, [vp, vp, vp, vp, sz]),
        "tfhe_b200_key_file_write": (i32, [C.c_char_p, vp, vp, vp, u32]),
        "tfhe_b200_key_file_info": (i32, [C.c_char_p, vp, vp, vp, vp]),
        "tfhe_b200_key_file_read": (i32, [C.c_char_p, vp, vp]),
        "tfhe_b200_key_file_last_error": (C.c_char_p, []),
        "tfhe_b200_load_key_file": (i32, [vp, C.c_char_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)   # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


EXPORTED_SYMBOLS = [
    "tfhe_b200_create", "tfhe_b200_destroy", "tfhe_b200_last_error", "tfhe_b200_num_devices", "tfhe_b200_version",
    "tfhe_b200_load_key", "tfhe_b200_load_key_device", "tfhe_b200_set_mode", "tfhe_b200_gate_batch", "tfhe_b200_gate_batch_ops",
    "tfhe_b200_bootstrap_batch", "tfhe_b200_bootstrap_no_keyswitch_batch", "tfhe_b200_blind_rotate_batch",
    "tfhe_b200_keyswitch_batch", "tfhe_b200_blind_rotate_extract_batch", "tfhe_b200_not_batch", "tfhe_b200_gate_batch_device",
    "tfhe_b200_bootstrap_batch_device", "tfhe_b200_blind_rotate_batch_device", "tfhe_b200_keyswitch_batch_device",
    "tfhe_b200_stream", "tfhe_b200_sync", "tfhe_b200_track_margin", "tfhe_b200_max_round_margin", "tfhe_b200_launch_count",
    "tfhe_b200_set_tuning", "tfhe_b200_measure_fp64_tflops", "tfhe_b200_last_kernel_ms",
    "tfhe_b200_load_reencryption_key", "tfhe_b200_reencrypt_batch",
    "tfhe_b200_keygen", "tfhe_b200_decomposition_offset", "tfhe_b200_lut_bootstrap_batch", "tfhe_b200_lut_generate",
    "tfhe_b200_lut_bootstrap_many_batch",
    "tfhe_b200_circuit_create", "tfhe_b200_circuit_destroy", "tfhe_b200_circuit_plan", "tfhe_b200_circuit_info", "tfhe_b200_circuit_run",
    "tfhe_b200_key_file_write", "tfhe_b200_key_file_info", "tfhe_b200_key_file_read", "tfhe_b200_key_file_last_error",
    "tfhe_b200_load_key_file",
]


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def _u32(a, width=None):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    if width is not None:
        a = a.reshape(-1, width)
    return a


@dataclass
class CloudKey:
    """key.CloudKey (key.zig:61-65): everything the evaluator needs, in the reference's layouts."""
    bootstrapping_key: np.ndarray            # f64 [n][2L][2][N]
    key_switching_key: np.ndarray | None     # u32 [N*t*base][n+1]
    decomposition_offset: int
    blind_rotate_testvec: np.ndarray | None = None   # u32 [2][N]; None = key.genTestvec default

    # flat cloud-key file (include/tfhe_b200.h "flat cloud-key file"); host only, no device needed
    def save(self, path, params: "str | Params"):
        key_file_write(path, params, self.bootstrapping_key, self.key_switching_key, self.decomposition_offset)

    @staticmethod
    def load(path) -> "CloudKey":
        return key_file_read(path)[1]


def _kf_check(lib, rc):
    if rc != 0:
        raise TfheB200Error(rc, (lib.tfhe_b200_key_file_last_error() or b"").decode())


def key_file_write(path, params: "str | Params", bsk, ksk, offset: int):
    lib = load_library()
    p = PARAM_SETS[params] if isinstance(params, str) else params
    cp = _Params(p.n, p.N, p.L, p.bgbit, p.basebit, p.iks_t)
    bsk = np.ascontiguousarray(bsk, dtype=np.float64)
    if bsk.size != p.n * 2 * p.L * 2 * N:
        raise ValueError("bootstrapping key has the wrong size")
    if ksk is not None:
        ksk = _u32(ksk)
        if ksk.size != N * p.iks_t * (1 << p.basebit) * (p.n + 1):
            raise ValueError("key-switching key has the wrong size")
    _kf_check(lib, lib.tfhe_b200_key_file_write(os.fsencode(path), C.byref(cp), _ptr(bsk), _ptr(ksk), int(offset) & 0xFFFFFFFF))


def key_file_info(path):
    """(Params, decomposition_offset, bsk_bytes, ksk_bytes) from the header alone"""
    lib = load_library()
    cp = _Params(); off = C.c_uint32(); nb = C.c_uint64(); nk = C.c_uint64()
    _kf_check(lib, lib.tfhe_b200_key_file_info(os.fsencode(path), C.byref(cp), C.byref(off), C.byref(nb), C.byref(nk)))
    name = next((k for k, q in PARAM_SETS.items()
                 if (q.n, q.N, q.L, q.bgbit, q.basebit, q.iks_t) == (cp.n, cp.N, cp.L, cp.bgbit, cp.basebit, cp.iks_t)), "custom")
    return Params(name, cp.n, cp.L, cp.bgbit, cp.basebit, cp.iks_t, cp.N), off.value, nb.value, nk.value


def key_file_read(path):
    """(Params, CloudKey) with both checksums verified"""
    lib = load_library()
    p, off, _, nk = key_file_info(path)
    bsk = np.empty((p.n, 2 * p.L, 2, N), np.float64)
    ksk = np.empty((N * p.iks_t * (1 << p.basebit), p.n + 1), np.uint32) if nk else None
    _kf_check(lib, lib.tfhe_b200_key_file_read(os.fsencode(path), _ptr(bsk), _ptr(ksk)))
    return p, CloudKey(bsk, ksk, off)


class Context:
    """One tfhe_b200_ctx.  `devices`: CUDA ordinals owned by this context (keys replicated on each)."""

    def __init__(self, params: str | Params = "128", devices=None):
        self.lib = load_library()
        self.params = PARAM_SETS[params] if isinstance(params, str) else params
        p = self.params
        cp = _Params(p.n, p.N, p.L, p.bgbit, p.basebit, p.iks_t)
        devices = [0] if devices is None else list(devices)
        ids = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        rc = self.lib.tfhe_b200_create(C.byref(cp), ids, len(devices), C.byref(h))
        if rc != 0:
            raise TfheB200Error(rc, "tfhe_b200_create failed (is a B200 / sm_100 device visible?)")
        self.h = h
        self.devices = devices
        self.n = p.n

    # -- plumbing
    def _check(self, rc):
        if rc != 0:
            msg = (self.lib.tfhe_b200_last_error(self.h) or b"").decode()
            raise (NotImplementedOnDevice if rc == 5 else TfheB200Error)(rc, msg)

    def close(self):
        if getattr(self, "h", None):
            self.lib.tfhe_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- keys
    def load_key(self, bsk, ksk, offset):
        bsk = np.ascontiguousarray(bsk, dtype=np.float64)
        p = self.params
        assert bsk.size == p.n * 2 * p.L * 2 * N, "bootstrapping key has the wrong size"
        stride = 0
        if ksk is not None:
            ksk = _u32(ksk)
            assert ksk.size == N * p.iks_t * (1 << p.basebit) * (p.n + 1), "key-switching key has the wrong size"
            stride = (p.n + 1) * 4
        self._check(self.lib.tfhe_b200_load_key(self.h, _ptr(bsk), _ptr(ksk), stride, int(offset) & 0xFFFFFFFF))

    def load_cloud_key(self, ck: CloudKey):
        self.load_key(ck.bootstrapping_key, ck.key_switching_key, ck.decomposition_offset)

    def keygen(self, key_lv0, key_lv1, seed: int, ksk_alpha: float, bsk_alpha: float, export: bool = True):
        """Generate the cloud key on the device(s) from a host-held secret key (key.CloudKey.new, key.zig:70-77).
        Returns a CloudKey in the reference layouts when `export`, else None."""
        s0 = _u32(key_lv0); s1 = _u32(key_lv1)
        assert s0.shape == (self.n,) and s1.shape == (1024,)
        p = self.params
        bsk = np.empty((p.n, 2 * p.L, 2, 1024), np.float64) if export else None
        ksk = np.empty((1024 * p.iks_t * (1 << p.basebit), p.n + 1), np.uint32) if export else None
        self._check(self.lib.tfhe_b200_keygen(self.h, _ptr(s0), _ptr(s1), int(seed) & 0xFFFFFFFFFFFFFFFF, float(ksk_alpha), float(bsk_alpha),
                                              _ptr(bsk), _ptr(ksk)))
        return CloudKey(bsk, ksk, int(self.lib.tfhe_b200_decomposition_offset(self.h))) if export else None

    def load_key_file(self, path):
        """load_key straight from a flat cloud-key file (mmap + checksum + upload), see key_file_write"""
        self._check(self.lib.tfhe_b200_load_key_file(self.h, os.fsencode(path)))

    def load_key_device(self, dev: int, d_bsk: int, d_ksk: int | None, offset: int):
        self._check(self.lib.tfhe_b200_load_key_device(self.h, dev, _ptr(d_bsk), _ptr(d_ksk), int(offset) & 0xFFFFFFFF))

    def set_mode(self, mode: int):
        self._check(self.lib.tfhe_b200_set_mode(self.h, mode))

    def set_tuning(self, key: str, value: int):
        self._check(self.lib.tfhe_b200_set_tuning(self.h, key.encode(), int(value)))

    # -- hot path (host buffers)
    def gate_batch(self, op, a, b, out=None):
        w = self.n + 1
        a = _u32(a, w); b = _u32(b, w)
        assert a.shape == b.shape
        B = a.shape[0]
        if out is None:
            out = np.empty((B, w), np.uint32)
        if np.isscalar(op):
            self._check(self.lib.tfhe_b200_gate_batch(self.h, int(op), _ptr(a), _ptr(b), _ptr(out), B))
        else:
            ops = np.ascontiguousarray(op, dtype=np.int32)
            assert ops.shape == (B,)
            self._check(self.lib.tfhe_b200_gate_batch_ops(self.h, _ptr(ops), _ptr(a), _ptr(b), _ptr(out), B))
        return out

    def bootstrap_batch(self, ct, testvec=None, tv_per_item=False, out=None):
        w = self.n + 1
        ct = _u32(ct, w); B = ct.shape[0]
        if out is None:
            out = np.empty((B, w), np.uint32)
        tv = _u32(testvec) if testvec is not None else None
        self._check(self.lib.tfhe_b200_bootstrap_batch(self.h, _ptr(ct), _ptr(out), B, _ptr(tv), 1 if tv_per_item else 0))
        return out

    def lut_bootstrap_batch(self, ct, tables, per_item=False):
        """tables: torus values [B][m] (per_item) or [m]; test vectors are generated on the device"""
        w = self.n + 1
        ct = _u32(ct, w); B = ct.shape[0]
        tables = np.ascontiguousarray(tables, dtype=np.uint32)
        m = tables.shape[-1]
        assert tables.shape == ((B, m) if per_item else (m,)), tables.shape
        out = np.empty((B, w), np.uint32)
        self._check(self.lib.tfhe_b200_lut_bootstrap_batch(self.h, _ptr(ct), _ptr(out), B, _ptr(tables), m, 1 if per_item else 0))
        return out

    def lut_bootstrap_many_batch(self, ct, tables):
        """several functions from one blind rotation: tables [k][m] torus values (k a power of two, k * 2m <= N) -> [k][B][n+1]"""
        w = self.n + 1
        ct = _u32(ct, w); B = ct.shape[0]
        tables = np.ascontiguousarray(tables, dtype=np.uint32)
        k, m = tables.shape
        out = np.empty((k, B, w), np.uint32)
        self._check(self.lib.tfhe_b200_lut_bootstrap_many_batch(self.h, _ptr(ct), _ptr(out), B, _ptr(tables), k, m))
        return out

    def lut_generate(self, table):
        table = np.ascontiguousarray(table, dtype=np.uint32)
        out = np.empty((2, 1024), np.uint32)
        self._check(self.lib.tfhe_b200_lut_generate(self.h, _ptr(table), table.shape[0], _ptr(out)))
        return out

    def bootstrap_no_keyswitch_batch(self, ct):
        w = self.n + 1
        ct = _u32(ct, w); B = ct.shape[0]
        out = np.empty((B, w), np.uint32)
        self._check(self.lib.tfhe_b200_bootstrap_no_keyswitch_batch(self.h, _ptr(ct), _ptr(out), B))
        return out

    def blind_rotate_batch(self, ct, testvec=None, tv_per_item=False):
        ct = _u32(ct, self.n + 1); B = ct.shape[0]
        out = np.empty((B, 2, N), np.uint32)
        tv = _u32(testvec) if testvec is not None else None
        self._check(self.lib.tfhe_b200_blind_rotate_batch(self.h, _ptr(ct), _ptr(out), B, _ptr(tv), 1 if tv_per_item else 0))
        return out

    def blind_rotate_extract_batch(self, ct):
        ct = _u32(ct, self.n + 1); B = ct.shape[0]
        out = np.empty((B, N + 1), np.uint32)
        self._check(self.lib.tfhe_b200_blind_rotate_extract_batch(self.h, _ptr(ct), _ptr(out), B))
        return out

    def keyswitch_batch(self, lv1):
        lv1 = _u32(lv1, N + 1); B = lv1.shape[0]
        out = np.empty((B, self.n + 1), np.uint32)
        self._check(self.lib.tfhe_b200_keyswitch_batch(self.h, _ptr(lv1), _ptr(out), B))
        return out

    def load_reencryption_key(self, key, basebit=None, t=None):
        """proxy_reenc.ProxyReencryptionKey (proxy_reenc.zig:123-252): u32 [n*t*base][n+1]"""
        p = self.params
        basebit = p.basebit if basebit is None else basebit
        t = p.iks_t if t is None else t
        key = _u32(key)
        assert key.size == p.n * t * (1 << basebit) * (p.n + 1), "re-encryption key has the wrong size"
        self._check(self.lib.tfhe_b200_load_reencryption_key(self.h, _ptr(key), basebit, t))

    def reencrypt_batch(self, ct):
        """proxy_reenc.reencryptTLWELv0 (proxy_reenc.zig:267-306) over a batch"""
        ct = _u32(ct, self.n + 1)
        out = np.empty_like(ct)
        self._check(self.lib.tfhe_b200_reencrypt_batch(self.h, _ptr(ct), _ptr(out), ct.shape[0]))
        return out

    def not_batch(self, a):
        a = _u32(a, self.n + 1)
        out = np.empty_like(a)
        self._check(self.lib.tfhe_b200_not_batch(self.h, _ptr(a), _ptr(out), a.shape[0]))
        return out

    # -- hot path (device pointers as ints, asynchronous on tfhe_b200_stream(dev))
    def gate_batch_device(self, dev, op, d_ops, d_a, d_b, d_out, B):
        self._check(self.lib.tfhe_b200_gate_batch_device(self.h, dev, int(op), _ptr(d_ops), _ptr(d_a), _ptr(d_b), _ptr(d_out), B))

    def bootstrap_batch_device(self, dev, d_in, d_out, B, d_tv=None, tv_per_item=False):
        self._check(self.lib.tfhe_b200_bootstrap_batch_device(self.h, dev, _ptr(d_in), _ptr(d_out), B, _ptr(d_tv), 1 if tv_per_item else 0))

    def blind_rotate_batch_device(self, dev, d_in, d_trlwe, B, d_tv=None, tv_per_item=False):
        self._check(self.lib.tfhe_b200_blind_rotate_batch_device(self.h, dev, _ptr(d_in), _ptr(d_trlwe), B, _ptr(d_tv), 1 if tv_per_item else 0))

    def keyswitch_batch_device(self, dev, d_lv1, d_lv0, B):
        self._check(self.lib.tfhe_b200_keyswitch_batch_device(self.h, dev, _ptr(d_lv1), _ptr(d_lv0), B))

    def stream(self, dev=0) -> int:
        return int(self.lib.tfhe_b200_stream(self.h, dev) or 0)

    def sync(self):
        self._check(self.lib.tfhe_b200_sync(self.h))

    # -- instrumentation
    def track_margin(self, enable=True):
        self._check(self.lib.tfhe_b200_track_margin(self.h, 1 if enable else 0))

    def max_round_margin(self, reset=True) -> float:
        return float(self.lib.tfhe_b200_max_round_margin(self.h, 1 if reset else 0))

    def launch_count(self) -> int:
        return int(self.lib.tfhe_b200_launch_count(self.h))

    def last_kernel_ms(self, dev=0, which=0) -> float:
        return float(self.lib.tfhe_b200_last_kernel_ms(self.h, dev, which))

    def measure_fp64_tflops(self, dev=0) -> float:
        return float(self.lib.tfhe_b200_measure_fp64_tflops(self.h, dev))



WIRE_NOT = 0x80000000
WIRE_TRUE, WIRE_FALSE = 0x7FFFFFFE, 0x7FFFFFFD    # Gates.constant(true / false) as circuit wires (gates.zig:144-151)

GATE_NODE = np.dtype([("op", np.int32), ("a", np.uint32), ("b", np.uint32)])


def circuit_plan(gates, n_inputs: int, outputs):
    """Host-only validation + levelisation of a netlist (tfhe_b200_circuit_plan; needs no GPU).
    Returns (n_levels, max_level_width, level of every gate); raises TfheB200Error on an invalid netlist."""
    lib = load_library()
    g = np.zeros(len(gates), GATE_NODE)
    for k, (op, a, b) in enumerate(gates):
        g[k] = (int(op), int(a), int(b))
    outs = np.ascontiguousarray(outputs, dtype=np.uint32)
    lv, wd = C.c_size_t(), C.c_size_t()
    gl = np.zeros(len(g), np.uint32)
    rc = lib.tfhe_b200_circuit_plan(_ptr(g), len(g), int(n_inputs), _ptr(outs), len(outs), C.byref(lv), C.byref(wd), _ptr(gl))
    if rc:
        raise TfheB200Error(rc, "invalid circuit")
    return lv.value, wd.value, gl


class Circuit:
    """A gate netlist compiled for a Context (tfhe_b200_circuit_*): levelised once, then every run evaluates all
    `instances` independent input sets level by level on the GPU -- the batched form of a chain of Gates.* calls
    such as examples/add_two_numbers.zig:24-73.

    gates: sequence of (op, a, b) in topological order; wire ids 0..n_inputs-1 are inputs, n_inputs+g the output of
    gate g; `wire | WIRE_NOT` feeds the Gates.notGate of a wire (free)."""

    def __init__(self, ctx: "Context", gates, n_inputs: int, outputs):
        self.ctx = ctx
        g = np.zeros(len(gates), GATE_NODE)
        for k, (op, a, b) in enumerate(gates):
            g[k] = (int(op), int(a), int(b))
        self.n_inputs = int(n_inputs)
        self.outputs = np.ascontiguousarray(outputs, dtype=np.uint32)
        h = C.c_void_p()
        ctx._check(ctx.lib.tfhe_b200_circuit_create(ctx.h, _ptr(g), len(g), self.n_inputs, _ptr(self.outputs), len(self.outputs), C.byref(h)))
        self.h = h
        lv, wd, ng = C.c_size_t(), C.c_size_t(), C.c_size_t()
        ctx._check(ctx.lib.tfhe_b200_circuit_info(self.h, C.byref(lv), C.byref(wd), C.byref(ng)))
        self.levels, self.max_width, self.n_gates = lv.value, wd.value, ng.value

    def run(self, inputs):
        """inputs: u32 [n_inputs][instances][n+1] -> outputs u32 [n_outputs][instances][n+1]"""
        w = self.ctx.n + 1
        x = np.ascontiguousarray(inputs, dtype=np.uint32)
        assert x.ndim == 3 and x.shape[0] == self.n_inputs and x.shape[2] == w, x.shape
        inst = x.shape[1]
        out = np.empty((len(self.outputs), inst, w), np.uint32)
        self.ctx._check(self.ctx.lib.tfhe_b200_circuit_run(self.ctx.h, self.h, _ptr(x), _ptr(out), inst))
        return out

    def close(self):
        if self.h is not None and self.ctx.h is not None:
            self.ctx.lib.tfhe_b200_circuit_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GpuBootstrap:
    """Bootstrap strategy backed by the device path; slots in where VanillaBootstrap does
    (bootstrap/vanilla.zig:25-75; the `Bootstrap` trait of bootstrap.zig:30-47)."""

    def __init__(self, ctx: Context):
        self.ctx = ctx

    def bootstrap(self, ctxt, cloud_key: CloudKey | None = None):
        """vanilla.zig:38-52 over a batch (a single ciphertext is a batch of one)."""
        one = np.asarray(ctxt).ndim == 1
        out = self.ctx.bootstrap_batch(ctxt, None if cloud_key is None else cloud_key.blind_rotate_testvec)
        return out[0] if one else out

    def bootstrap_without_key_switch(self, ctxt, cloud_key: CloudKey | None = None):
        """vanilla.zig:58-69"""
        one = np.asarray(ctxt).ndim == 1
        out = self.ctx.bootstrap_no_keyswitch_batch(ctxt)
        return out[0] if one else out

    def name(self) -> str:
        return "b200"


class Gates:
    """gates.Gates (gates.zig:25-151) on batches: every method takes [B][n+1] arrays (or a single
    ciphertext) and returns the same shape."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self.bootstrap = GpuBootstrap(ctx)

    def bootstrap_strategy(self) -> str:
        return self.bootstrap.name()

    def _gate(self, op, a, b):
        one = np.asarray(a).ndim == 1
        out = self.ctx.gate_batch(op, a, b)
        return out[0] if one else out

    def nand(self, a, b): return self._gate(NAND, a, b)
    def or_(self, a, b): return self._gate(OR, a, b)
    def and_(self, a, b): return self._gate(AND, a, b)
    def xor(self, a, b): return self._gate(XOR, a, b)
    def xnor(self, a, b): return self._gate(XNOR, a, b)
    def nor(self, a, b): return self._gate(NOR, a, b)
    def and_ny(self, a, b): return self._gate(ANDNY, a, b)
    def and_yn(self, a, b): return self._gate(ANDYN, a, b)
    def or_ny(self, a, b): return self._gate(ORNY, a, b)
    def or_yn(self, a, b): return self._gate(ORYN, a, b)

    def not_(self, a):
        """gates.zig:131-134 (no bootstrap)"""
        one = np.asarray(a).ndim == 1
        out = self.ctx.not_batch(a)
        return out[0] if one else out

    def copy(self, a):
        return np.array(a, dtype=np.uint32, copy=True)

    def constant(self, value: bool):
        """gates.zig:144-151 (false is 1 - 2^29, reference quirk kept)"""
        out = np.zeros(self.ctx.n + 1, np.uint32)
        out[-1] = 0x20000000 if value else (1 - 0x20000000) & 0xFFFFFFFF
        return out

    def mux_naive(self, a, b, c):
        """gates.zig:124-129: (a AND b) OR ((NOT a) AND c)"""
        a_and_b = self.and_(a, b)
        nand_a_c = self.and_(self.not_(a), c)
        return self.or_(a_and_b, nand_a_c)


def _batch(op):
    def f(ctx: Context, inputs, cloud_key: CloudKey | None = None):
        """gates.batch* (gates.zig:244-295): inputs = sequence of (a, b) ciphertext pairs."""
        a = np.stack([np.asarray(x[0], dtype=np.uint32) for x in inputs])
        b = np.stack([np.asarray(x[1], dtype=np.uint32) for x in inputs])
        return ctx.gate_batch(op, a, b)
    return f


batch_nand, batch_and, batch_or = _batch(NAND), _batch(AND), _batch(OR)
batch_xor, batch_nor, batch_xnor = _batch(XOR), _batch(NOR), _batch(XNOR)
