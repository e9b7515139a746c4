"""Level-batched ripple-carry adder (BASELINE config 3; workload from examples/add_two_numbers.zig:24-73)."""
import numpy as np
import pytest

from conftest import keys_for
from oracle import oracle as O


class PlainCtx:
    """stand-in context whose 'ciphertexts' carry the plaintext bit in the body slot: checks the circuit wiring,
    level count and gate count on CPU without any cryptography"""
    n = 3

    def __init__(self):
        self.calls = []

    def gate_batch(self, op, a, b):
        ops = np.full(len(a), op, np.int32) if np.isscalar(op) else np.asarray(op)
        self.calls.append(len(a))
        x, y = a[:, -1], b[:, -1]
        r = np.where(ops == 2, x & y, np.where(ops == 3, x ^ y, x | y))
        out = np.zeros_like(a)
        out[:, -1] = r
        return out


def _plain(bits):
    c = np.zeros(bits.shape + (4,), np.uint32)
    c[..., -1] = bits
    return c


def test_adder_wiring_levels_and_gate_count():
    from tfhe_b200 import circuits
    rng = np.random.default_rng(0)
    W, B = 16, 37
    x = rng.integers(0, 2**16, B); y = rng.integers(0, 2**16, B)
    x[0], y[0] = 402, 304                                   # add_two_numbers.zig:103-104
    ctx = PlainCtx()
    sums, carry, gates, levels = circuits.ripple_carry_add(ctx, _plain(circuits.to_bits(x, W)), _plain(circuits.to_bits(y, W)), _plain(np.zeros(B, np.uint8)))
    total = circuits.from_bits(sums[..., -1]) + (carry[:, -1].astype(np.uint64) << np.uint64(W))
    assert (total == x + y).all() and total[0] == 706
    assert levels == 33 and gates == 80 * B
    assert ctx.calls == [2 * W * B] + [2 * B, B] * W


@pytest.mark.gpu
def test_adder_on_gpu_matches_plaintext_and_oracle():
    import tfhe_b200
    from tfhe_b200 import circuits
    orc = O.Oracle("128"); keys = keys_for("128")
    ctx = tfhe_b200.Context("128", devices=[0])
    try:
        ctx.load_key(keys.bsk, keys.ksk, keys.offset)
        rng = np.random.default_rng(42)
        W, B = 16, 64
        x = rng.integers(0, 2**16, B); y = rng.integers(0, 2**16, B)
        x[0], y[0] = 402, 304
        enc = lambda bits, seed: np.stack([orc.encrypt_bools(bits[i], keys, seed + i) for i in range(W)])
        ca, cb = enc(circuits.to_bits(x, W), 1000), enc(circuits.to_bits(y, W), 2000)
        cin = orc.encrypt_bools(np.zeros(B, np.uint8), keys, 3000)
        sums, carry, gates, levels = circuits.ripple_carry_add(ctx, ca, cb, cin)
        dec = np.stack([orc.decrypt_bools(sums[i], keys) for i in range(W)])
        total = circuits.from_bits(dec) + (orc.decrypt_bools(carry, keys).astype(np.uint64) << np.uint64(W))
        assert (total == x + y).all() and total[0] == 706
        assert gates == 80 * B and levels == 33
        # instance 0, first two bit positions, gate by gate against the reference evaluation order (fullAdder)
        c = cin[0]
        for i in range(2):
            axb = orc.gate(O.XOR, ca[i, 0], cb[i, 0], keys); ab = orc.gate(O.AND, ca[i, 0], cb[i, 0], keys)
            t = orc.gate(O.AND, axb, c, keys); s = orc.gate(O.XOR, axb, c, keys)
            c = orc.gate(O.OR, ab, t, keys)
            assert (sums[i, 0] == s).all()
    finally:
        ctx.close()


@pytest.mark.gpu
def test_adder_full_size_1024_instances():
    """BASELINE config 3 at full size: 1,024 parallel 16-bit additions = 81,920 bootstrapped gates, 33 levels"""
    import tfhe_b200
    from tfhe_b200 import circuits, hostkeys as HK
    params = tfhe_b200.PARAM_SETS["128"]
    sk, ck = HK.gen_cloud_key(params, seed=1)
    ctx = tfhe_b200.Context(params, devices=[0])
    try:
        ctx.load_cloud_key(ck)
        rng = np.random.default_rng(7)
        W, B = 16, 1024
        x = rng.integers(0, 2**16, B); y = rng.integers(0, 2**16, B)
        x[0], y[0] = 402, 304
        enc = lambda bits: np.stack([HK.encrypt_bools(bits[i], params, sk, rng) for i in range(W)])
        sums, carry, gates, levels = circuits.ripple_carry_add(ctx, enc(circuits.to_bits(x, W)), enc(circuits.to_bits(y, W)),
                                                               HK.encrypt_bools(np.zeros(B, np.uint8), params, sk, rng))
        dec = np.stack([HK.decrypt_bools(sums[i], sk) for i in range(W)])
        total = circuits.from_bits(dec) + (HK.decrypt_bools(carry, sk).astype(np.uint64) << np.uint64(W))
        assert (total == x + y).all() and gates == 81920 and levels == 33
    finally:
        ctx.close()


@pytest.mark.gpu
def test_device_resident_adder_equals_host_level_version():
    import torch
    import tfhe_b200
    from tfhe_b200 import circuits, hostkeys as HK
    params = tfhe_b200.PARAM_SETS["128"]
    sk, ck = HK.gen_cloud_key(params, seed=2)
    ctx = tfhe_b200.Context(params, devices=[0])
    try:
        ctx.load_cloud_key(ck)
        rng = np.random.default_rng(9)
        W, B = 16, 200
        x = rng.integers(0, 2**16, B); y = rng.integers(0, 2**16, B)
        enc = lambda bits: np.stack([HK.encrypt_bools(bits[i], params, sk, rng) for i in range(W)])
        ca, cb = enc(circuits.to_bits(x, W)), enc(circuits.to_bits(y, W))
        cin = HK.encrypt_bools(np.zeros(B, np.uint8), params, sk, rng)
        sums_h, carry_h, _, _ = circuits.ripple_carry_add(ctx, ca, cb, cin)
        t = lambda a: torch.from_numpy(a.view(np.int32)).cuda()
        sums_d, carry_d = circuits.ripple_carry_add_device(ctx, t(ca), t(cb), t(cin))
        assert (sums_d.cpu().numpy().view(np.uint32) == sums_h).all()
        assert (carry_d.cpu().numpy().view(np.uint32) == carry_h).all()
        dec = np.stack([HK.decrypt_bools(sums_h[i], sk) for i in range(W)])
        total = circuits.from_bits(dec) + (HK.decrypt_bools(carry_h, sk).astype(np.uint64) << np.uint64(W))
        assert (total == x + y).all()
    finally:
        ctx.close()


# ---- native circuit executor (tfhe_b200_circuit_*) ----------------------------------------------------------
def _eval_netlist_plain(gates, n_inputs, outputs, in_bits):
    """reference semantics of a netlist on plaintext bits (gates.zig:48-121 truth tables, conftest.TRUTH)"""
    from conftest import TRUTH
    NOT = 0x80000000
    wires = [np.asarray(b, np.uint8) for b in in_bits]
    ref = lambda w: (1 - wires[w & ~NOT]) if (w & NOT) else wires[w & ~NOT]
    for op, a, b in gates:
        wires.append(np.asarray(TRUTH[op](ref(a), ref(b)), np.uint8))
    return [ref(o) for o in outputs]


def test_adder_netlist_is_the_reference_full_adder_chain():
    from tfhe_b200 import circuits
    W = 16
    gates, n_in, outs = circuits.ripple_carry_netlist(W)
    assert len(gates) == 5 * W and n_in == 2 * W + 1 and len(outs) == W + 1      # add_two_numbers.zig:24-73
    for k, (_, a, b) in enumerate(gates):
        assert max(a, b) < n_in + k                                              # topological
    rng = np.random.default_rng(5)
    B = 50
    x = rng.integers(0, 2**16, B); y = rng.integers(0, 2**16, B); cin = rng.integers(0, 2, B)
    x[0], y[0], cin[0] = 402, 304, 0
    bits = list(circuits.to_bits(x, W)) + list(circuits.to_bits(y, W)) + [cin.astype(np.uint8)]
    o = _eval_netlist_plain(gates, n_in, outs, bits)
    total = circuits.from_bits(np.stack(o[:W])) + (o[W].astype(np.uint64) << np.uint64(W))
    assert (total == x + y + cin).all() and total[0] == 706


def test_circuit_plan_levels_without_a_gpu():
    """tfhe_b200_circuit_plan is the host-only half of circuit_create: the 16-bit adder has 1 + 2 * 16 levels, the first one
    32 gates wide; invalid netlists are rejected with the same codes"""
    import tfhe_b200
    from tfhe_b200 import circuits
    gates, n_in, outs = circuits.ripple_carry_netlist(16)
    levels, width, gl = tfhe_b200.circuit_plan(gates, n_in, outs)
    assert (levels, width) == (33, 32)
    assert (gl[0::5] == 1).all() and (gl[1::5] == 1).all()                     # every a^b and a&b: level 1
    assert list(gl[2::5]) == [2 * i + 2 for i in range(16)] == list(gl[3::5])  # (a^b)&c and the sum bit of bit i
    assert list(gl[4::5]) == [2 * i + 3 for i in range(16)]                    # carry out of bit i
    assert tfhe_b200.circuit_plan([], 3, [0, 2 | tfhe_b200.WIRE_NOT])[:2] == (0, 0)   # wires only
    for bad_gates, bad_outs in (([(0, 0, 3)], [2]), ([(17, 0, 1)], [2]), ([(0, 0, 1)], [9])):
        with pytest.raises(tfhe_b200.TfheB200Error):
            tfhe_b200.circuit_plan(bad_gates, 2, bad_outs)


@pytest.mark.gpu
def test_native_circuit_equals_level_batched_python_and_oracle():
    import tfhe_b200
    from tfhe_b200 import circuits
    orc = O.Oracle("128"); keys = keys_for("128")
    ctx = tfhe_b200.Context("128", devices=[0])
    try:
        ctx.load_key(keys.bsk, keys.ksk, keys.offset)
        rng = np.random.default_rng(7)
        W, B = 4, 37                                          # ragged instance count
        x = rng.integers(0, 2**W, B); y = rng.integers(0, 2**W, B)
        enc = lambda bits, seed: np.stack([orc.encrypt_bools(bits[i], keys, seed + i) for i in range(W)])
        ca, cb = enc(circuits.to_bits(x, W), 100), enc(circuits.to_bits(y, W), 200)
        cin = orc.encrypt_bools(np.zeros(B, np.uint8), keys, 300)
        s_py, c_py, gates, levels = circuits.ripple_carry_add(ctx, ca, cb, cin)
        s_nat, c_nat, circ = circuits.ripple_carry_add_native(ctx, ca, cb, cin)
        assert circ.levels == 1 + 2 * W == levels and circ.n_gates == 5 * W and circ.max_width == 2 * W
        assert (s_nat == s_py).all() and (c_nat == c_py).all()            # every ciphertext word
        dec = np.stack([orc.decrypt_bools(s_nat[i], keys) for i in range(W)])
        total = circuits.from_bits(dec) + (orc.decrypt_bools(c_nat, keys).astype(np.uint64) << np.uint64(W))
        assert (total == x + y).all()
        # graph replay (second run, same size), then eager levels: identical bits
        s2, c2, _ = circuits.ripple_carry_add_native(ctx, ca, cb, cin, circ)
        ctx.set_tuning("circuit_graph", 0)
        s3, c3, _ = circuits.ripple_carry_add_native(ctx, ca, cb, cin, circ)
        assert (s2 == s_nat).all() and (c2 == c_nat).all() and (s3 == s_nat).all() and (c3 == c_nat).all()
        ctx.set_tuning("circuit_graph", 1)
        # a different instance count re-captures
        s4, c4, _ = circuits.ripple_carry_add_native(ctx, ca[:, :5], cb[:, :5], cin[:5], circ)
        assert (s4 == s_nat[:, :5]).all() and (c4 == c_nat[:5]).all()
        circ.close()
    finally:
        ctx.close()


@pytest.mark.gpu
def test_native_circuit_not_folding_and_all_gates_vs_oracle():
    """NOT wires are folded into the consumer's linear part (gates.zig:131-133 then 48-121): every opcode with
    plain and negated operands, and a negated circuit output, word for word against the oracle"""
    import tfhe_b200
    orc = O.Oracle("128"); keys = keys_for("128")
    ctx = tfhe_b200.Context("128", devices=[0])
    try:
        ctx.load_key(keys.bsk, keys.ksk, keys.offset)
        NOT = tfhe_b200.WIRE_NOT
        B = 3
        a_bits = np.array([0, 1, 1], np.uint8); b_bits = np.array([1, 0, 1], np.uint8)
        ca = orc.encrypt_bools(a_bits, keys, 11); cb = orc.encrypt_bools(b_bits, keys, 12)
        gates = [(op, 0, 1) for op in range(10)] + [(op, 0 | NOT, 1) for op in range(10)] + [(O.AND, 0, 1 | NOT)]
        gates.append((O.OR, 2 + 2, 2 + 12 | NOT))                 # second level: OR(AND(a,b), NOT ANDNY... any wires)
        outs = list(range(2, 2 + len(gates))) + [(2 + 3) | NOT, 0 | NOT]
        circ = tfhe_b200.Circuit(ctx, gates, 2, outs)
        assert circ.levels == 2
        got = circ.run(np.stack([ca, cb]))
        neg = lambda c: (0 - c.astype(np.int64)).astype(np.uint32)
        wires = [ca, cb]
        ref = lambda w: neg(wires[w & ~NOT]) if (w & NOT) else wires[w & ~NOT]
        for op, a, b in gates:
            wires.append(orc.gate_batch(op, ref(a), ref(b), keys))
        for k, o in enumerate(outs):
            assert (got[k] == ref(o)).all(), f"output {k}"
        circ.close()
        # error behaviour: forward reference, bad opcode, bad output wire
        for bad in ([(0, 0, 3)], [(17, 0, 1)]):
            with pytest.raises(tfhe_b200.TfheB200Error):
                tfhe_b200.Circuit(ctx, bad, 2, [2])
        with pytest.raises(tfhe_b200.TfheB200Error):
            tfhe_b200.Circuit(ctx, [(0, 0, 1)], 2, [9])
    finally:
        ctx.close()


@pytest.mark.gpu
def test_mux_naive_circuit_matches_the_reference_composition():
    """gates.zig:124-129 and its truth-table test (gates.zig:513-544) as one two-level circuit over all 8 input rows"""
    import tfhe_b200
    from tfhe_b200 import circuits
    orc = O.Oracle("128"); keys = keys_for("128")
    ctx = tfhe_b200.Context("128", devices=[0])
    try:
        ctx.load_key(keys.bsk, keys.ksk, keys.offset)
        rows = np.array([[a, b, c] for a in (0, 1) for b in (0, 1) for c in (0, 1)], np.uint8)
        ca = orc.encrypt_bools(rows[:, 0], keys, 41); cb = orc.encrypt_bools(rows[:, 1], keys, 42); cc = orc.encrypt_bools(rows[:, 2], keys, 43)
        out, circ = circuits.mux_naive_batch(ctx, ca, cb, cc)
        assert circ.levels == 2 and circ.n_gates == 3
        assert (orc.decrypt_bools(out, keys) == np.where(rows[:, 0] == 1, rows[:, 1], rows[:, 2])).all()
        ref = np.stack([orc.gate(O.OR, orc.gate(O.AND, ca[i], cb[i], keys), orc.gate(O.AND, orc.gate_not(ca[i]), cc[i], keys), keys) for i in range(8)])
        assert (out == ref).all()
        circ.close()
    finally:
        ctx.close()


@pytest.mark.gpu
def test_circuit_graph_is_recaptured_after_rekey_and_mode_switch():
    """the captured level graph bakes in key pointers, mode and tuning.  Same circuit, same instance
    count: run, load another cloud key, run again (oracle under the NEW key), then FAST -> EXACT -> FAST."""
    import tfhe_b200
    from tfhe_b200 import circuits
    orc = O.Oracle("128"); k1 = keys_for("128"); k2 = keys_for("128", seed=2)
    ctx = tfhe_b200.Context("128", devices=[0])
    try:
        gates, n_in, outs = circuits.mux_naive_netlist()
        rows = np.array([[a, b, c] for a in (0, 1) for b in (0, 1) for c in (0, 1)], np.uint8)
        want = np.where(rows[:, 0] == 1, rows[:, 1], rows[:, 2])

        def oracle_mux(k, ca, cb, cc):
            return orc.gate_batch(O.OR, orc.gate_batch(O.AND, ca, cb, k), orc.gate_batch(O.AND, np.stack([orc.gate_not(x) for x in ca]), cc, k), k)

        ctx.load_key(k1.bsk, k1.ksk, k1.offset)
        circ = tfhe_b200.Circuit(ctx, gates, n_in, outs)
        in1 = [orc.encrypt_bools(rows[:, j], k1, 41 + j) for j in range(3)]
        out1 = circ.run(np.stack(in1))[0]
        assert (out1 == oracle_mux(k1, *in1)).all()
        ctx.load_key(k2.bsk, k2.ksk, k2.offset)          # frees and reallocates every key buffer
        in2 = [orc.encrypt_bools(rows[:, j], k2, 51 + j) for j in range(3)]
        out2 = circ.run(np.stack(in2))[0]
        assert (orc.decrypt_bools(out2, k2) == want).all()
        assert (out2 == oracle_mux(k2, *in2)).all()
        launches = ctx.launch_count()
        ctx.set_mode(tfhe_b200.MODE_EXACT)               # the exact kernel must actually run: different kernel, same bits here
        out3 = circ.run(np.stack(in2))[0]
        assert (out3 == out2).all() and ctx.launch_count() > launches
        ctx.set_mode(tfhe_b200.MODE_FAST)
        ctx.set_tuning("kct", 2)
        assert (circ.run(np.stack(in2))[0] == out2).all()
        ctx.set_tuning("kct", 0)
        circ.close()
    finally:
        ctx.close()


def test_circuit_plan_accepts_constant_wires():
    import tfhe_b200
    T, F, NOT = tfhe_b200.WIRE_TRUE, tfhe_b200.WIRE_FALSE, tfhe_b200.WIRE_NOT
    levels, width, gl = tfhe_b200.circuit_plan([(O.AND, 0, T), (O.OR, 2, F | NOT), (O.NAND, T, F)], 2, [3, T, F | NOT])
    assert (levels, width) == (2, 2) and list(gl) == [1, 2, 1]
    assert tfhe_b200.circuit_plan([(O.XOR, T, F)], 0, [0])[:2] == (1, 1)            # no inputs at all: wire 0 is the gate


@pytest.mark.gpu
def test_constant_wires_in_circuits_match_gates_constant():
    """Gates.constant (src/gates.zig:144-151) as circuit wires: (0, 2^29) and the reference's (0, 1 - 2^29); as gate operands,
    negated, and as outputs; word for word against the oracle's gate_constant + gates"""
    import tfhe_b200
    orc = O.Oracle("128"); keys = keys_for("128")
    ctx = tfhe_b200.Context("128", devices=[0])
    try:
        ctx.load_key(keys.bsk, keys.ksk, keys.offset)
        T, F, NOT = tfhe_b200.WIRE_TRUE, tfhe_b200.WIRE_FALSE, tfhe_b200.WIRE_NOT
        bits = np.array([0, 1, 1, 0, 1], np.uint8)
        ca = orc.encrypt_bools(bits, keys, 61)
        gates = [(O.AND, 0, T), (O.OR, 0, F), (O.XOR, 0, T | NOT), (O.NAND, T, F), (O.ANDNY, 1, T)]
        outs = [1, 2, 3, 4, 5, T, F, F | NOT]
        circ = tfhe_b200.Circuit(ctx, gates, 1, outs)
        assert circ.levels == 2
        got = circ.run(ca[None])
        B = len(bits)
        ct = np.stack([orc.gate_constant(True)] * B); cf = np.stack([orc.gate_constant(False)] * B)
        neg = lambda c: (0 - c.astype(np.int64)).astype(np.uint32)
        w1 = orc.gate_batch(O.AND, ca, ct, keys)
        want = [w1, orc.gate_batch(O.OR, ca, cf, keys), orc.gate_batch(O.XOR, ca, neg(ct), keys), orc.gate_batch(O.NAND, ct, cf, keys),
                orc.gate_batch(O.ANDNY, w1, ct, keys), ct, cf, neg(cf)]
        for k, w in enumerate(want):
            assert (got[k] == w).all(), f"output {k}"
        assert (orc.decrypt_bools(got[0], keys) == bits).all() and (orc.decrypt_bools(got[3], keys) == 1).all()
        assert got[6][0, -1] == 0xE0000001 and got[5][0, -1] == 0x20000000
        circ.close()
    finally:
        ctx.close()
