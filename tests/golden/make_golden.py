"""Regenerates tests/golden/oracle_128.json from the CPU oracle (seeded).  The reference has no golden
vectors of its own (and cannot be built here), so these fixtures are regression guards on the ORACLE'S
bits: they catch compiler-flag drift (FMA contraction, reassociation) and accidental edits, and they are
what the GPU tests compare against when no live oracle run is wanted.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def build():
    orc = O.Oracle("128")
    keys = orc.keygen(seed=1)
    a = np.array([0, 0, 1, 1], np.uint8); b = np.array([0, 1, 0, 1], np.uint8)
    ca = orc.encrypt_bools(a, keys, seed=11); cb = orc.encrypt_bools(b, keys, seed=12)
    lin = np.stack([orc.gate_linear(O.NAND, ca[i], cb[i]) for i in range(4)])
    tr = orc.blind_rotate_batch(lin, keys)
    lv1 = np.stack([orc.sample_extract_index(tr[i], 0) for i in range(4)])
    out = orc.keyswitch_batch(lv1, keys)
    x = np.arange(1024, dtype=np.uint32) * np.uint32(2654435761)
    g = {
        "params": "128", "key_seed": 1, "input_seeds": [11, 12],
        "offset": int(keys.offset),
        "sha256": {"s0": sha(keys.s0), "s1": sha(keys.s1), "bsk": sha(keys.bsk), "ksk": sha(keys.ksk), "ca": sha(ca), "cb": sha(cb),
                   "lin_nand": sha(lin), "blind_rotate": sha(tr), "extract": sha(lv1), "nand_out": sha(out),
                   "ifft1024": sha(O.ifft1024(x)), "fft1024_roundtrip": sha(O.fft1024(O.ifft1024(x)))},
        "nand_out_row0": out[0].tolist(),
        "nand_bits": orc.decrypt_bools(out, keys).tolist(),
        "ifft1024_head": O.ifft1024(x)[:8].tolist(),
    }
    return g


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_128.json")
    json.dump(build(), open(path, "w"), indent=1)
    print("wrote", path)
