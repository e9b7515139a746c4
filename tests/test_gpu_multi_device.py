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
