"""Oracle and CUDA path against dumps of a REAL zig-tfhe build (tools/zig_golden/README.md).

Skipped unless tests/golden/zig/manifest.txt exists: no Zig toolchain was available where this repo was written, so the
dumps are not committed; producing them is one `zig build-exe` + one run on any machine with Zig 0.15.1."""
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as O

D = os.path.join(ROOT, "tests", "golden", "zig")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(D, "manifest.txt")), reason="no Zig golden dumps (tools/zig_golden/README.md)")


def _load():
    n, N, L, bgbit, basebit, t, offset, G, R = (int(x) for x in open(os.path.join(D, "manifest.txt")).read().split())
    name = {(550, 7): "80", (630, 8): "110", (700, 9): "128"}[(n, t)]
    orc = O.Oracle(name)
    assert (orc.n, orc.L, orc.bgbit, orc.basebit, orc.iks_t) == (n, L, bgbit, basebit, t) and N == 1024
    rd = lambda f, dt, shape: np.fromfile(os.path.join(D, f), dtype=dt).reshape(shape)
    sec = rd("secret.bin", np.uint32, (n + N,))
    keys = O.Keys(sec[:n].copy(), sec[n:].copy(), rd("bsk.bin", np.float64, (n, 2 * L, 2, N)), rd("ksk.bin", np.uint32, (N * t * (1 << basebit), n + 1)),
                  offset, None)
    z = {"ops": rd("gate_ops.bin", np.int32, (G,)), "a": rd("gate_a.bin", np.uint32, (G, n + 1)), "b": rd("gate_b.bin", np.uint32, (G, n + 1)),
         "out": rd("gate_out.bin", np.uint32, (G, n + 1)), "bits": rd("gate_bits.bin", np.uint8, (G, 3)),
         "rin": rd("rot_in.bin", np.uint32, (R, n + 1)), "rtr": rd("rot_trlwe.bin", np.uint32, (R, 2, N)),
         "rl1": rd("rot_lv1.bin", np.uint32, (R, N + 1)), "rl0": rd("rot_lv0.bin", np.uint32, (R, n + 1))}
    return name, orc, keys, z


def test_zig_dump_is_self_consistent():
    from conftest import TRUTH
    name, orc, keys, z = _load()
    assert keys.offset == int(O.lib().orc_decomposition_offset(orc._pp))
    for op, (a, b, r) in zip(z["ops"], z["bits"]):
        assert r == TRUTH[int(op)](int(a), int(b))
    assert (orc.decrypt_bools(z["out"], keys) == z["bits"][:, 2]).all()
    # the dumped bootstrapping key is a forward transform of integer polynomials (exact round trip, src/trlwe.zig:111-132)
    back, margin = O.fft1024(keys.bsk[0, 0, 0], with_margin=True)
    assert margin < 1e-3 and (O.ifft1024(back) == keys.bsk[0, 0, 0]).all()


def test_oracle_reproduces_zig_bits():
    name, orc, keys, z = _load()
    assert (orc.gate_batch(z["ops"], z["a"], z["b"], keys) == z["out"]).all()
    assert (orc.blind_rotate_batch(z["rin"], keys) == z["rtr"]).all()
    assert (np.stack([orc.sample_extract_index(t, 0) for t in z["rtr"]]) == z["rl1"]).all()
    assert (orc.keyswitch_batch(z["rl1"], keys) == z["rl0"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fast", "exact"])
def test_gpu_reproduces_zig_bits(mode):
    import tfhe_b200
    name, orc, keys, z = _load()
    c = tfhe_b200.Context(name, devices=[0])
    try:
        c.load_key(keys.bsk, keys.ksk, keys.offset)
        c.set_mode(tfhe_b200.MODE_EXACT if mode == "exact" else tfhe_b200.MODE_FAST)
        assert (c.gate_batch(z["ops"], z["a"], z["b"]) == z["out"]).all()
        assert (c.blind_rotate_batch(z["rin"]) == z["rtr"]).all()
        assert (c.blind_rotate_extract_batch(z["rin"]) == z["rl1"]).all()
        for tc in (-1, 1):
            c.set_tuning("ks_tc", tc)
            assert (c.keyswitch_batch(z["rl1"]) == z["rl0"]).all()
    finally:
        c.close()
