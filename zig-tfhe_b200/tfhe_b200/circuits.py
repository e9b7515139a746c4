"""Level-batched gate circuits on top of the batch gate API.

Workload definition from the reference's examples/add_two_numbers.zig:24-73 (fullAdder + ripple-carry
`add`): per bit  a^b, a&b, (a^b)&c, (a^b)^c, (a&b)|((a^b)&c)  -- 5 bootstrapped gates, 80 for 16 bits,
evaluated strictly one after another by the reference.  Here every dependency level of the circuit is ONE
batched call over all independent additions: level 0 holds all W a^b and W a&b gates,
then two levels per bit; 1 + 2W levels, 5W gates per addition.  Instances are independent, so a multi-GPU
context shards them with no cross-device traffic.
"""
from __future__ import annotations

import numpy as np

from . import AND, OR, XOR, Context


def to_bits(values, width: int) -> np.ndarray:
    """bit_utils.convert (bit_utils.zig:16-29): LSB first, shape [width][len(values)]"""
    v = np.asarray(values, dtype=np.uint64)
    return np.stack([((v >> np.uint64(i)) & np.uint64(1)).astype(np.uint8) for i in range(width)])


def from_bits(bits) -> np.ndarray:
    bits = np.asarray(bits, dtype=np.uint64)
    return sum(bits[i] << np.uint64(i) for i in range(bits.shape[0]))


def ripple_carry_add(ctx: Context, a_bits: np.ndarray, b_bits: np.ndarray, cin: np.ndarray):
    """a_bits, b_bits: [W][B][n+1] ciphertexts (LSB first), cin: [B][n+1].
    Returns (sum_bits [W][B][n+1], carry [B][n+1], gates evaluated, levels)."""
    W, B, w = a_bits.shape
    assert b_bits.shape == a_bits.shape and cin.shape == (B, w)
    # level 0: every a^b and a&b of every bit position at once
    a_flat = a_bits.reshape(W * B, w)
    b_flat = b_bits.reshape(W * B, w)
    ops0 = np.concatenate([np.full(W * B, XOR, np.int32), np.full(W * B, AND, np.int32)])
    lvl0 = ctx.gate_batch(ops0, np.concatenate([a_flat, a_flat]), np.concatenate([b_flat, b_flat]))
    a_xor_b = lvl0[: W * B].reshape(W, B, w)
    a_and_b = lvl0[W * B:].reshape(W, B, w)
    gates, levels = 2 * W * B, 1
    carry = np.ascontiguousarray(cin)
    sums = np.empty_like(a_bits)
    ops1 = np.concatenate([np.full(B, AND, np.int32), np.full(B, XOR, np.int32)])
    for i in range(W):
        # level 2i+1: (a^b)&c and (a^b)^c
        x = np.concatenate([a_xor_b[i], a_xor_b[i]])
        lvl = ctx.gate_batch(ops1, x, np.concatenate([carry, carry]))
        sums[i] = lvl[B:]
        # level 2i+2: carry = (a&b) | ((a^b)&c)
        carry = ctx.gate_batch(OR, a_and_b[i], lvl[:B])
        gates += 3 * B
        levels += 2
    return sums, carry, gates, levels
