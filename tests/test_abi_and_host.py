"""CPU-side checks of the boundary: the CUDA library loads and exports every symbol the header declares,
it refuses to run without a GPU (no CPU fallback), and the host-side key code agrees with the oracle."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, keys_for
from oracle import oracle as O


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "tfhe_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(tfhe_b200_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    import ctypes
    import tfhe_b200
    lib = tfhe_b200.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/tfhe_b200.h but not exported"
    assert set(tfhe_b200.EXPORTED_SYMBOLS) == set(declared)
    assert b"sm_100a" in ctypes.cast(lib.tfhe_b200_version(), ctypes.c_char_p).value


def test_no_cpu_fallback():
    """on a box without an sm_100 device the product fails loudly instead of computing on the CPU"""
    import torch
    import tfhe_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(tfhe_b200.TfheB200Error) as e:
        tfhe_b200.Context("128")
    assert e.value.code == 2


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "zig-tfhe_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".zig")):
                src = open(os.path.join(d, f), errors="ignore").read()
                assert "tfhe_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f


def test_param_tables_agree():
    import tfhe_b200
    for name, p in tfhe_b200.PARAM_SETS.items():
        o = O.Oracle(name)
        assert (p.n, p.L, p.bgbit, p.basebit, p.iks_t) == (o.n, o.L, o.bgbit, o.basebit, o.iks_t)


def test_hostkeys_match_reference_semantics():
    """keys made by the product's host-side keygen work under the oracle's (reference) evaluator"""
    import tfhe_b200
    from tfhe_b200 import hostkeys as HK
    params = tfhe_b200.PARAM_SETS["128"]
    sk, ck = HK.gen_cloud_key(params, seed=3)
    orc = O.Oracle("128")
    assert ck.decomposition_offset == 0x82080000
    keys = O.Keys(sk.key_lv0, sk.key_lv1, ck.bootstrapping_key, ck.key_switching_key, ck.decomposition_offset, ck.blind_rotate_testvec)
    rng = np.random.default_rng(0)
    a = rng.integers(0, 2, 8).astype(np.uint8); b = rng.integers(0, 2, 8).astype(np.uint8)
    ca = HK.encrypt_bools(a, params, sk, rng); cb = HK.encrypt_bools(b, params, sk, rng)
    assert (HK.decrypt_bools(ca, sk) == a).all() and (orc.decrypt_bools(ca, keys) == a).all()
    out = orc.gate_batch(O.XOR, ca, cb, keys)
    assert (HK.decrypt_bools(out, sk) == a ^ b).all()
    x = rng.integers(0, 2**32, 1024, dtype=np.uint32)
    ref = O.ifft1024(x)
    assert np.abs(HK.spectrum(x) - ref).max() < 1e-12 * np.abs(ref).max()
    assert (HK.f64_to_torus(np.array([0.125, -0.125, 0.25, -1e-20])) == np.array([0x20000000, 0xE0000000, 0x40000000, 0], np.uint32)).all()


def test_gates_mirror_constant_and_shapes():
    import tfhe_b200

    class FakeCtx:
        n = 700
    g = tfhe_b200.Gates.__new__(tfhe_b200.Gates)
    g.ctx = FakeCtx()
    assert g.constant(True)[-1] == 0x20000000 and g.constant(False)[-1] == 0xE0000001     # gates.zig:146-147
    assert g.constant(True).shape == (701,)
