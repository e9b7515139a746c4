"""Host-side mirror of the reference's `lut` module (src/lut.zig, src/lut/{encoder,generator,lookup_table}.zig)
plus the `bootstrapLut` entry point its documentation promises (src/lut.zig:42) but never defines.

The generator is tiny host code (it builds one TRLWE test vector); the bootstrap itself runs on the device through
tfhe_b200_bootstrap_batch with that test vector (trgsw.blindRotateWithTestvec, src/trgsw.zig:336-400)."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import Context
from .hostkeys import f64_to_torus

N = 1024


class Encoder:
    """lut/encoder.zig:12-114: message m in [0, modulus) <-> torus value m * scale, scale = 1/(2*modulus)"""

    def __init__(self, message_modulus: int, scale: float | None = None):
        self.message_modulus = int(message_modulus)
        self.scale = 1.0 / (2.0 * self.message_modulus) if scale is None else float(scale)   # encoder.zig:35, 44-49

    def encode(self, message: int) -> int:                                                   # encoder.zig:66-73
        return int(f64_to_torus(np.float64((message % self.message_modulus) * self.scale)))

    def encode_with_scale(self, message: int, scale: float) -> int:                          # encoder.zig:83-87
        return int(f64_to_torus(np.float64((message % self.message_modulus) * float(scale))))

    def decode(self, value: int) -> int:                                                     # encoder.zig:96-105
        f = float(int(value) & 0xFFFFFFFF) / 4294967296.0
        return int(f / self.scale + 0.5) % self.message_modulus

    def decode_bool(self, value: int) -> bool:                                               # encoder.zig:111-113
        return self.decode(value) != 0


@dataclass
class LookupTable:
    """lut/lookup_table.zig:16-58: a TRLWE (a = 0) whose b polynomial is the rotated, sign-folded table"""
    poly: np.ndarray = field(default_factory=lambda: np.zeros((2, N), np.uint32))

    def is_empty(self) -> bool:
        return not self.poly.any()

    def clear(self):
        self.poly[:] = 0

    def copy_from(self, other: "LookupTable"):
        self.poly[:] = other.poly


def _div_round(a: int, b: int) -> int:      # generator.zig:253-255
    return (a + b // 2) // b


class Generator:
    """lut/generator.zig:15-250"""

    def __init__(self, message_modulus: int, scale: float | None = None):
        self.encoder = Encoder(message_modulus, scale)
        self.poly_degree = N
        self.lookup_table_size = N

    def generate_lookup_table(self, f) -> LookupTable:
        """generator.zig:65-135: f maps a message index to a message"""
        return self._generate([self.encoder.encode(f(x)) for x in range(self.encoder.message_modulus)])

    def generate_lookup_table_full(self, f) -> LookupTable:
        """generator.zig:150-191: f returns raw torus values"""
        return self._generate([int(f(x)) & 0xFFFFFFFF for x in range(self.encoder.message_modulus)])

    def generate_lookup_table_custom(self, f, message_modulus: int, scale: float) -> LookupTable:
        """generator.zig:202-212: the same generator with a temporary Encoder.withScale(message_modulus, scale)"""
        return Generator(message_modulus, scale).generate_lookup_table(f)

    def function_table(self, f) -> np.ndarray:
        """`message_modulus` torus words Encoder.encode(f(x)): the compact form tfhe_b200_lut_bootstrap_batch /
        tfhe_b200_lut_generate take (the LookupTable itself is then built on the device, generator.zig:150-191)"""
        return np.array([self.encoder.encode(f(x)) for x in range(self.encoder.message_modulus)], np.uint32)

    def _generate(self, encoded) -> LookupTable:
        m, size = self.encoder.message_modulus, self.lookup_table_size
        raw = np.zeros(size, np.uint32)
        for x in range(m):                                        # generator.zig:95-110
            raw[_div_round(x * size, m): _div_round((x + 1) * size, m)] = encoded[x]
        offset = _div_round(size, 2 * m)                          # generator.zig:113
        rotated = np.roll(raw, -offset)                           # generator.zig:120-123
        rotated[size - offset:] = (0 - rotated[size - offset:].astype(np.int64)).astype(np.uint32)   # :126-128
        lut = LookupTable()
        lut.poly[1] = rotated                                     # generator.zig:131-134 (a stays 0)
        return lut

    def mod_switch(self, x: int) -> int:                          # generator.zig:223-227
        scaled = (float(x) / float(0xFFFFFFFF)) * self.lookup_table_size
        return int(scaled + 0.5) % self.lookup_table_size


def bootstrap_lut(ctx: Context, ciphertexts, lut: LookupTable | np.ndarray, per_item: bool = False):
    """`VanillaBootstrap.bootstrapLut` of src/lut.zig:42 (documented, not implemented upstream), batched:
    blindRotateWithTestvec -> sampleExtractIndex(., 0) -> identityKeySwitching on the device."""
    tv = lut.poly if isinstance(lut, LookupTable) else np.asarray(lut, dtype=np.uint32)
    one = np.asarray(ciphertexts).ndim == 1
    out = ctx.bootstrap_batch(ciphertexts, tv, tv_per_item=per_item)
    return out[0] if one else out
