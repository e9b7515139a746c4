"""Host-side lut mirror (encoder / generator / LookupTable) against the oracle's restatement, and bootstrapLut on GPU."""
import numpy as np
import pytest

from conftest import keys_for
from oracle import oracle as O


def test_lut_generator_matches_oracle():
    from tfhe_b200 import lut
    orc = O.Oracle("128")
    for m in (2, 3, 4, 8, 16, 32):
        gen = lut.Generator(m)
        for name, f in (("id", lambda x: x), ("inc", lambda x: (x + 1) % m), ("sq", lambda x: (x * x) % m), ("not", lambda x: m - 1 - x)):
            table = np.array([f(x) for x in range(m)], np.uint32)
            assert (gen.generate_lookup_table(f).poly == orc.lut_generate(table, m)).all(), (m, name)
        enc = lut.Encoder(m)
        for x in range(m):
            assert enc.encode(x) == O.lut_encode(x, m) and enc.decode(enc.encode(x)) == x == O.lut_decode(enc.encode(x), m)
    t = lut.LookupTable()
    assert t.is_empty()                                   # lookup_table.zig "lookup table creation"
    t.poly[1, 0] = 42
    u = lut.LookupTable(); u.copy_from(t)
    assert not u.is_empty() and u.poly[1, 0] == 42        # "lookup table copy"
    u.clear()
    assert u.is_empty()
    assert lut._div_round(7, 2) == 4 and lut._div_round(1024, 8) == 128   # generator.zig "div round"


@pytest.mark.gpu
def test_bootstrap_lut_on_gpu():
    import tfhe_b200
    from tfhe_b200 import lut
    orc = O.Oracle("128"); k = keys_for("128")
    c = tfhe_b200.Context("128", devices=[0])
    try:
        c.load_key(k.bsk, k.ksk, k.offset)
        m = 4
        msgs = np.arange(12, dtype=np.uint32) % m
        ct = orc.encrypt_lwe_messages(msgs, m, k, seed=77)
        gen = lut.Generator(m)
        table = lut.Generator(m).generate_lookup_table(lambda x: (3 * x + 1) % m)
        out = lut.bootstrap_lut(c, ct, table)
        assert (orc.decrypt_lwe_messages(out, m, k) == (3 * msgs + 1) % m).all()
        assert (out == orc.bootstrap_batch(ct, k, table.poly)).all()
        assert (lut.bootstrap_lut(c, ct[0], table) == out[0]).all()
    finally:
        c.close()


def test_reference_generator_and_encoder_tests_mirrored():
    """lut/generator.zig:259-345 and lut/encoder.zig tests, same cases, plus what they leave unchecked"""
    from tfhe_b200 import lut
    g = lut.Generator(2)
    assert (g.encoder.message_modulus, g.poly_degree, g.lookup_table_size) == (2, 1024, 1024)      # "generator creation"
    for f in (lambda x: x, lambda x: 1 - x, lambda x: 1):                                          # identity / not / constant
        assert not g.generate_lookup_table(f).is_empty()
    assert not lut.Generator(4).generate_lookup_table(lambda x: (x + 1) % 4).is_empty()            # "4bit function"
    half = lut.Generator(2, 0.5)                                                                   # "custom scale"
    t = half.generate_lookup_table(lambda x: x)
    assert not t.is_empty()
    # message 1 at scale 0.5 is torus 1/2: windows [256, 768) hold 0x80000000, the rest 0 (and -0 = 0 in the tail)
    assert set(np.unique(t.poly[1]).tolist()) == {0, 0x80000000} and not t.poly[0].any()
    for x in (0, 0xFFFFFFFF // 2, 0xFFFFFFFF):                                                     # "mod switch"
        assert 0 <= g.mod_switch(x) < g.lookup_table_size
    assert g.mod_switch(0) == 0 and g.mod_switch(0xFFFFFFFF // 2) == 512
    # generateLookupTableCustom (generator.zig:202-212) = a generator with Encoder.withScale(modulus, scale)
    c = g.generate_lookup_table_custom(lambda x: (x + 1) % 4, 4, 0.125)
    assert (c.poly == lut.Generator(4).generate_lookup_table(lambda x: (x + 1) % 4).poly).all()   # 0.125 = default scale of modulus 4
    assert (g.generate_lookup_table_custom(lambda x: x, 2, 0.5).poly == t.poly).all()
    # encodeWithScale (encoder.zig:83-87)
    e = lut.Encoder(4)
    assert e.encode_with_scale(3, 0.125) == e.encode(3) == 0x60000000 and e.encode_with_scale(5, 0.25) == 0x40000000
    assert e.decode_bool(e.encode(0)) is False and e.decode_bool(e.encode(2)) is True
    # the compact table form the device generator consumes
    assert (lut.Generator(4).function_table(lambda x: 3 - x) == [0x60000000, 0x40000000, 0x20000000, 0]).all()


@pytest.mark.gpu
def test_device_lut_generator_and_table_bootstrap():
    """tfhe_b200_lut_generate == the host mirror == the oracle for every modulus (incl. non powers of two), and
    tfhe_b200_lut_bootstrap_batch with per-item function tables == bootstrap with the host-built per-item test vectors"""
    import tfhe_b200
    from tfhe_b200 import lut
    orc = O.Oracle("128"); k = keys_for("128")
    c = tfhe_b200.Context("128", devices=[0])
    try:
        c.load_key(k.bsk, k.ksk, k.offset)
        rng = np.random.default_rng(4)
        for m in (1, 2, 3, 4, 5, 7, 8, 16, 31, 32, 100, 256, 1024):
            table = rng.integers(0, 2**32, m, dtype=np.uint32)
            gen = lut.Generator(m)
            host = gen.generate_lookup_table_full(lambda x: int(table[x])).poly
            assert (c.lut_generate(table) == host).all(), m
        m = 4
        enc = lut.Encoder(m)
        B = 9
        msgs = np.arange(B, dtype=np.uint32) % m
        ct = orc.encrypt_lwe_messages(msgs, m, k, seed=78)
        fs = [lambda x, a=a: (a * x + 1) % m for a in range(B)]              # a different function per item
        tables = np.array([[enc.encode(f(x)) for x in range(m)] for f in fs], np.uint32)
        tvs = np.stack([lut.Generator(m).generate_lookup_table(f).poly for f in fs])
        out = c.lut_bootstrap_batch(ct, tables, per_item=True)
        assert (out == c.bootstrap_batch(ct, tvs, tv_per_item=True)).all()
        want = np.array([fs[i](int(msgs[i])) for i in range(B)])
        assert (orc.decrypt_lwe_messages(out, m, k) == want).all()
        one = c.lut_bootstrap_batch(ct, tables[2])                             # one table shared by the batch
        assert (one == c.bootstrap_batch(ct, tvs[2])).all()
    finally:
        c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,m,k", [("128", 4, 4), ("128", 4, 2), ("uint1", 2, 8)])
def test_many_function_bootstrap(name, m, k):
    """several functions from one blind rotation (tfhe_b200_lut_bootstrap_many_batch): every output decodes f_j(message), and
    equals, word for word, the oracle's blindRotateWithTestvec on the interleaved test vector fed with the input rounded to
    the coarse modulus-switch grid, sampleExtractIndex(., j), identityKeySwitching (src/trgsw.zig:336-400, 471-502,
    src/trlwe.zig:146-162)"""
    import tfhe_b200
    from conftest import keys_for
    orc = O.Oracle(name); keys = keys_for(name)
    c = tfhe_b200.Context(name, devices=[0])
    try:
        c.load_key(keys.bsk, keys.ksk, keys.offset)
        if name != "128":
            c.set_mode(tfhe_b200.MODE_EXACT)
        B = 300
        rng = np.random.default_rng(4)
        msgs = rng.integers(0, m, B).astype(np.uint32)
        ct = orc.encrypt_lwe_messages(msgs, m, keys, seed=9)
        funcs = [np.array([(x * (j + 1) + j) % m for x in range(m)], np.uint32) for j in range(k)]
        tables = np.stack([np.array([O.lut_encode(int(v), m) for v in f], np.uint32) for f in funcs])
        got = c.lut_bootstrap_many_batch(ct, tables)
        assert got.shape == (k, B, orc.n + 1)
        for j in range(k):
            assert (orc.decrypt_lwe_messages(got[j], m, keys) == funcs[j][msgs]).all(), f"function {j}"
        # oracle emulation on a sample: pre-round the input to multiples of k * 2^21, interleave the test vectors
        shift = k.bit_length() - 1
        sel = np.arange(0, B, 37)
        pre = (((ct[sel].astype(np.uint64) + (1 << (20 + shift))) >> (21 + shift)) << (21 + shift)).astype(np.uint32)
        tv = np.zeros((2, 1024), np.uint32)
        for j in range(k):
            tj = orc.lut_generate(funcs[j], m)
            tv[1, j::k] = tj[1, 0::k]
        tr = orc.blind_rotate_batch(pre, keys, tv)
        for j in range(k):
            lv1 = np.stack([orc.sample_extract_index(t, j) for t in tr])
            assert (got[j][sel] == orc.keyswitch_batch(lv1, keys)).all(), f"function {j} differs from the oracle"
        with pytest.raises(tfhe_b200.TfheB200Error):
            c.lut_bootstrap_many_batch(ct, np.zeros((3, m), np.uint32))      # not a power of two
    finally:
        c.close()
