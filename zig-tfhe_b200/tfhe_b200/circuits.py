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

from . import AND, OR, WIRE_NOT, XOR, Circuit, Context


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


def ripple_carry_add_device(ctx: Context, d_a, d_b, d_cin, dev: int = 0):
    """Same circuit with every ciphertext resident on the GPU between levels:
    d_a, d_b: torch int32 tensors [W][B][n+1] on the context's device `dev`, d_cin: [B][n+1].
    All work (level gathers and the K1/K2 launches) is enqueued on the library's stream; nothing returns to the host
    until the caller reads the result.  Returns (sum_bits [W][B][n+1], carry [B][n+1]) as torch tensors."""
    import torch

    W, B, w = d_a.shape
    stream = torch.cuda.ExternalStream(ctx.stream(dev), device=d_a.device)
    with torch.cuda.stream(stream):
        a_flat = d_a.reshape(W * B, w).contiguous()
        b_flat = d_b.reshape(W * B, w).contiguous()
        a_xor_b = torch.empty_like(a_flat)
        a_and_b = torch.empty_like(a_flat)
        ctx.gate_batch_device(dev, XOR, None, a_flat.data_ptr(), b_flat.data_ptr(), a_xor_b.data_ptr(), W * B)
        ctx.gate_batch_device(dev, AND, None, a_flat.data_ptr(), b_flat.data_ptr(), a_and_b.data_ptr(), W * B)
        a_xor_b = a_xor_b.view(W, B, w)
        a_and_b = a_and_b.view(W, B, w)
        ops = torch.cat([torch.full((B,), AND, dtype=torch.int32), torch.full((B,), XOR, dtype=torch.int32)]).to(d_a.device)
        x2 = torch.empty((2 * B, w), dtype=torch.int32, device=d_a.device)
        c2 = torch.empty_like(x2)
        lvl = torch.empty_like(x2)
        sums = torch.empty_like(d_a)
        carry = d_cin.contiguous().clone()
        for i in range(W):
            x2[:B].copy_(a_xor_b[i]); x2[B:].copy_(a_xor_b[i])
            c2[:B].copy_(carry); c2[B:].copy_(carry)
            ctx.gate_batch_device(dev, 0, ops.data_ptr(), x2.data_ptr(), c2.data_ptr(), lvl.data_ptr(), 2 * B)
            sums[i].copy_(lvl[B:])
            ctx.gate_batch_device(dev, OR, None, a_and_b[i].data_ptr(), lvl.data_ptr(), carry.data_ptr(), B)
        # keep every temporary alive until the stream has consumed it
        stream.synchronize()
    return sums, carry


def ripple_carry_netlist(width: int):
    """The reference's adder as a netlist for Circuit (examples/add_two_numbers.zig:24-73, fullAdder evaluated bit by
    bit).  Inputs: a_0..a_{W-1}, b_0..b_{W-1}, cin (2W+1 wires); outputs: sum_0..sum_{W-1}, carry.
    Returns (gates, n_inputs, outputs); 5W gates, 1 + 2W levels."""
    W = width
    n_inputs = 2 * W + 1
    gates, outputs = [], []
    carry = 2 * W                       # cin
    for i in range(W):
        a, b = i, W + i
        g = n_inputs + len(gates)
        gates.append((XOR, a, b))       # g     = a ^ b
        gates.append((AND, a, b))       # g + 1 = a & b
        gates.append((AND, g, carry))   # g + 2 = (a ^ b) & c
        gates.append((XOR, g, carry))   # g + 3 = sum
        gates.append((OR, g + 1, g + 2))  # g + 4 = carry out
        outputs.append(g + 3)
        carry = g + 4
    outputs.append(carry)
    return gates, n_inputs, outputs


def ripple_carry_add_native(ctx: Context, a_bits: np.ndarray, b_bits: np.ndarray, cin: np.ndarray, circuit: Circuit | None = None):
    """Same result as ripple_carry_add through the native circuit executor (one C call, wires stay on the GPU, levels
    replayed as a CUDA graph).  Returns (sum_bits [W][B][n+1], carry [B][n+1], circuit)."""
    W, B, w = a_bits.shape
    if circuit is None:
        gates, n_in, outs = ripple_carry_netlist(W)
        circuit = Circuit(ctx, gates, n_in, outs)
    out = circuit.run(np.concatenate([a_bits, b_bits, cin[None]]))
    return out[:W], out[W], circuit


def mux_naive_netlist():
    """Gates.muxNaive (gates.zig:124-129): a ? b : c = (a & b) | (~a & c); inputs a, b, c; the NOT is folded into the AND."""
    return [(AND, 0, 1), (AND, 0 | WIRE_NOT, 2), (OR, 3, 4)], 3, [5]


def mux_naive_batch(ctx: Context, a: np.ndarray, b: np.ndarray, c: np.ndarray, circuit: Circuit | None = None):
    """`count` independent muxNaive evaluations in one circuit call (2 levels, 3 bootstraps each); a, b, c: [count][n+1]."""
    if circuit is None:
        gates, n_in, outs = mux_naive_netlist()
        circuit = Circuit(ctx, gates, n_in, outs)
    return circuit.run(np.stack([a, b, c]))[0], circuit
