"""Flat cloud-key file (include/tfhe_b200.h "flat cloud-key file").  The write/info/read calls are host
only, so they run without a GPU; the format is also restated here in numpy so that the documented layout and checksum are
what the library actually writes."""
import os
import struct

import numpy as np
import pytest

import tfhe_b200
from tfhe_b200 import Params

SMALL = Params("custom", 5, 2, 10, 2, 3)     # tiny but structurally complete: 160 KiB BSK, 288 KiB KSK
PRIME, BASIS, M64 = 0x100000001B3, 0xCBF29CE484222325, (1 << 64) - 1


def doc_checksum(buf: bytes) -> int:
    """the checksum as specified in the header comment, written independently of the C code"""
    n = len(buf)
    words = list(struct.unpack(f"<{n // 8}Q", buf[: n // 8 * 8]))
    if n % 8:
        words.append(int.from_bytes(buf[n // 8 * 8:], "little"))
    lane = [BASIS, BASIS + 1, BASIS + 2, BASIS + 3]
    for i, w in enumerate(words):
        lane[i & 3] = ((lane[i & 3] ^ w) * PRIME) & M64
    h = BASIS
    for l in lane:
        h = ((h ^ l) * PRIME) & M64
    return ((h ^ n) * PRIME) & M64


def small_key(seed=0, with_ksk=True):
    rng = np.random.default_rng(seed)
    p = SMALL
    bsk = rng.standard_normal((p.n, 2 * p.L, 2, 1024))
    ksk = rng.integers(0, 2**32, (1024 * p.iks_t * (1 << p.basebit), p.n + 1), dtype=np.uint64).astype(np.uint32) if with_ksk else None
    return bsk, ksk, 0x82080000


def test_round_trip_and_documented_layout(tmp_path):
    path = tmp_path / "cloud.key"
    bsk, ksk, off = small_key()
    tfhe_b200.key_file_write(path, SMALL, bsk, ksk, off)
    p, o, nb, nk = tfhe_b200.key_file_info(path)
    assert (p.n, p.N, p.L, p.bgbit, p.basebit, p.iks_t) == (5, 1024, 2, 10, 2, 3) and o == off
    assert nb == bsk.nbytes and nk == ksk.nbytes
    p2, ck = tfhe_b200.key_file_read(path)
    assert p2 == p and ck.decomposition_offset == off
    assert ck.bootstrapping_key.tobytes() == bsk.tobytes() and (ck.key_switching_key == ksk).all()

    raw = path.read_bytes()                       # the layout of the header comment, field by field
    assert raw[:8] == b"TFHEB2CK"
    version, hdr = struct.unpack_from("<II", raw, 8)
    assert (version, hdr) == (1, 4096)
    assert struct.unpack_from("<6i", raw, 16) == (5, 1024, 2, 10, 2, 3)
    assert struct.unpack_from("<II", raw, 40) == (off, 1)
    bo, bb, ko, kb, bsum, ksum, hsum = struct.unpack_from("<7Q", raw, 48)
    assert bo == 4096 and bb == bsk.nbytes and ko % 4096 == 0 and ko >= bo + bb and kb == ksk.nbytes
    assert len(raw) == ko + kb and raw[104:4096] == bytes(4096 - 104)
    assert raw[bo:bo + bb] == bsk.tobytes() and raw[ko:ko + kb] == ksk.tobytes()   # sections usable in place
    assert bsum == doc_checksum(bsk.tobytes()) and ksum == doc_checksum(ksk.tobytes()) and hsum == doc_checksum(raw[:96])
    assert not [f for f in os.listdir(tmp_path) if ".tmp." in f]                   # temporary renamed away


def test_checksum_of_odd_lengths():
    """the trailing partial word: exercised through the header (96 bytes) only in the file, so check the rule directly"""
    for n in (0, 1, 7, 8, 9, 31, 32, 33):
        buf = bytes(range(1, n + 1))
        a = doc_checksum(buf); b = doc_checksum(buf + b"\0")
        assert a != b                             # the length is part of the sum: zero padding is not invisible


def test_file_without_key_switching_key(tmp_path):
    """CloudKey.newNoKsk (key.zig:80-100)"""
    path = tmp_path / "noksk.key"
    bsk, _, off = small_key(1, with_ksk=False)
    tfhe_b200.CloudKey(bsk, None, off).save(path, SMALL)
    _, _, nb, nk = tfhe_b200.key_file_info(path)
    assert nb == bsk.nbytes and nk == 0
    ck = tfhe_b200.CloudKey.load(path)
    assert ck.key_switching_key is None and ck.bootstrapping_key.tobytes() == bsk.tobytes()
    lib = tfhe_b200.load_library()
    dummy = np.zeros(8, np.uint32)
    rc = lib.tfhe_b200_key_file_read(os.fsencode(path), None, dummy.ctypes.data)
    assert rc == 4 and b"no key-switching key" in lib.tfhe_b200_key_file_last_error()


def test_corruption_is_detected(tmp_path):
    path = tmp_path / "cloud.key"
    bsk, ksk, off = small_key(2)
    tfhe_b200.key_file_write(path, SMALL, bsk, ksk, off)
    good = path.read_bytes()

    def expect(mutated: bytes, text: str, code: int = 1, info_fails: bool = False):
        path.write_bytes(mutated)
        with pytest.raises(tfhe_b200.TfheB200Error) as e:
            tfhe_b200.key_file_read(path)
        assert e.value.code == code and text in str(e.value), str(e.value)
        if not info_fails:
            tfhe_b200.key_file_info(path)         # header intact: info still answers

    flip = lambda i: good[:i] + bytes([good[i] ^ 0x10]) + good[i + 1:]
    expect(flip(4096 + 12345), "bootstrapping-key checksum")
    ko = struct.unpack_from("<Q", good, 64)[0]
    expect(flip(ko + 999), "key-switching-key checksum")
    expect(flip(0), "bad magic", info_fails=True)
    expect(flip(8), "unknown version", info_fails=True)
    expect(flip(17), "header checksum", info_fails=True)          # n changed
    expect(good[:-1], "truncated", info_fails=True)
    expect(good[:50], "too short", code=6, info_fails=True)
    with pytest.raises(tfhe_b200.TfheB200Error) as e:
        tfhe_b200.key_file_info(tmp_path / "missing.key")
    assert e.value.code == 6 and "cannot open" in str(e.value)


def test_bad_arguments(tmp_path):
    bsk, ksk, off = small_key(3)
    with pytest.raises(ValueError):
        tfhe_b200.key_file_write(tmp_path / "x", SMALL, bsk[:-1], ksk, off)
    with pytest.raises(tfhe_b200.TfheB200Error) as e:           # N must be 1024 (all 11 reference sets)
        tfhe_b200.key_file_write(tmp_path / "x", Params("bad", 5, 2, 10, 2, 3, N=512), bsk[:, :, :, :512].repeat(2, 3), ksk, off)
    assert e.value.code == 1
    with pytest.raises(tfhe_b200.TfheB200Error) as e:
        tfhe_b200.key_file_write(tmp_path / "no_such_dir" / "x", SMALL, bsk, ksk, off)
    assert e.value.code == 6
    assert not (tmp_path / "x").exists()


@pytest.mark.gpu
def test_load_key_file_equals_load_key(tmp_path):
    """generate on the device -> file -> a second context loads the file: same ciphertext bits as the generating context
    and as the oracle under the key read back from the file; a context of another parameter set refuses the file"""
    from conftest import keys_for
    from oracle import oracle as O
    orc = O.Oracle("128"); ref = keys_for("128")
    path = tmp_path / "cloud128.key"
    c1 = tfhe_b200.Context("128"); c2 = tfhe_b200.Context("128"); c3 = tfhe_b200.Context("110")
    try:
        ck = c1.keygen(ref.s0, ref.s1, seed=42, ksk_alpha=2.0e-5, bsk_alpha=2.0e-8)
        ck.save(path, "128")
        c2.load_key_file(path)
        back = tfhe_b200.CloudKey.load(path)
        keys = O.Keys(ref.s0, ref.s1, back.bootstrapping_key, back.key_switching_key, back.decomposition_offset, ref.testvec)
        rng = np.random.default_rng(5)
        B = 300
        a = rng.integers(0, 2, B).astype(np.uint8); b = rng.integers(0, 2, B).astype(np.uint8)
        ca = orc.encrypt_bools(a, keys, 7); cb = orc.encrypt_bools(b, keys, 8)
        out1 = c1.gate_batch(tfhe_b200.NAND, ca, cb); out2 = c2.gate_batch(tfhe_b200.NAND, ca, cb)
        assert (out1 == out2).all()
        assert (orc.decrypt_bools(out2, keys) == 1 - (a & b)).all()
        assert (out2[:3] == orc.gate_batch(O.NAND, ca[:3], cb[:3], keys)).all()
        with pytest.raises(tfhe_b200.TfheB200Error) as e:
            c3.load_key_file(path)
        assert e.value.code == 1 and "key is for n=700" in str(e.value)
        with pytest.raises(tfhe_b200.TfheB200Error) as e:
            c2.load_key_file(tmp_path / "missing.key")
        assert e.value.code == 6
        assert (c2.gate_batch(tfhe_b200.NAND, ca[:5], cb[:5]) == out2[:5]).all()       # a failed load leaves the loaded key alone
    finally:
        c1.close(); c2.close(); c3.close()
