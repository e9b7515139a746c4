"""One context owning several GPUs (the in-process form a Zig caller uses): contiguous batch split, keys
replicated, same bits as a single device.  Skipped when fewer than 2 GPUs are visible."""
import numpy as np
import pytest

from conftest import keys_for
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_two_device_context_matches_single_device():
    import torch
    import tfhe_b200
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    orc = O.Oracle("128"); k = keys_for("128")
    rng = np.random.default_rng(3)
    B = 301
    a = rng.integers(0, 2, B).astype(np.uint8); b = rng.integers(0, 2, B).astype(np.uint8)
    ca = orc.encrypt_bools(a, k, 5); cb = orc.encrypt_bools(b, k, 6)
    c2 = tfhe_b200.Context("128", devices=[0, 1]); c1 = tfhe_b200.Context("128", devices=[1])
    try:
        for c in (c1, c2):
            c.load_key(k.bsk, k.ksk, k.offset)
        out2 = c2.gate_batch(O.XOR, ca, cb)
        out1 = c1.gate_batch(O.XOR, ca, cb)
        assert (out1 == out2).all()
        assert (orc.decrypt_bools(out2, k) == a ^ b).all()
        assert (out2[::50] == orc.gate_batch(O.XOR, ca[::50], cb[::50], k)).all()
    finally:
        c1.close(); c2.close()


def test_circuit_shards_instances_over_devices_and_lanes():
    """tfhe_b200_circuit_run on a two-device context (4 lanes each): contiguous instance ranges per device and lane,
    no cross-device traffic, same bits as one device with one lane"""
    import torch
    import tfhe_b200
    from tfhe_b200 import circuits
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    orc = O.Oracle("128"); k = keys_for("128")
    rng = np.random.default_rng(11)
    W, B = 3, 29                                     # 29 instances over 2 devices x 4 lanes: ragged everywhere
    x = rng.integers(0, 2**W, B); y = rng.integers(0, 2**W, B)
    enc = lambda bits, seed: np.stack([orc.encrypt_bools(bits[i], k, seed + i) for i in range(W)])
    ca, cb = enc(circuits.to_bits(x, W), 10), enc(circuits.to_bits(y, W), 20)
    cin = orc.encrypt_bools(np.zeros(B, np.uint8), k, 30)
    c2 = tfhe_b200.Context("128", devices=[0, 1]); c1 = tfhe_b200.Context("128", devices=[0])
    try:
        c1.set_tuning("circuit_lanes", 1)
        for c in (c1, c2):
            c.load_key(k.bsk, k.ksk, k.offset)
        s2, carry2, q2 = circuits.ripple_carry_add_native(c2, ca, cb, cin)
        s1, carry1, q1 = circuits.ripple_carry_add_native(c1, ca, cb, cin)
        assert (s1 == s2).all() and (carry1 == carry2).all()
        dec = np.stack([orc.decrypt_bools(s2[i], k) for i in range(W)])
        total = circuits.from_bits(dec) + (orc.decrypt_bools(carry2, k).astype(np.uint64) << np.uint64(W))
        assert (total == x + y).all()
        q1.close(); q2.close()
    finally:
        c1.close(); c2.close()


def test_failure_on_one_device_of_a_multi_device_context():
    """a shard failing on device 1 while device 0 works: the call fails as a whole (err_mu-guarded message from the failing
    thread), nothing hangs, and the context computes correctly afterwards"""
    import torch
    import tfhe_b200
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    orc = O.Oracle("128"); keys = keys_for("128")
    c = tfhe_b200.Context("128", devices=[0, 1])
    try:
        c.load_key(keys.bsk, keys.ksk, keys.offset)
        bits = (np.arange(300) % 2).astype(np.uint8)
        ca = orc.encrypt_bools(bits, keys, 5); cb = orc.encrypt_bools(1 - bits, keys, 6)
        c.set_tuning("inject_fault", 2)
        with pytest.raises(tfhe_b200.TfheB200Error, match="injected fault on device 1"):
            c.gate_batch(O.OR, ca, cb)
        out = c.gate_batch(O.OR, ca, cb)
        assert (orc.decrypt_bools(out, keys) == 1).all()
        sel = np.array([0, 149, 150, 299])
        assert (out[sel] == orc.gate_batch(O.OR, ca[sel], cb[sel], keys)).all()
    finally:
        c.close()
