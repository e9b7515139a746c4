"""Writes the file set golden_dump.zig writes, but from the CPU oracle: a SELF-TEST of tests/test_zig_golden.py's loader
(formats, shapes, skip logic).  It pins nothing -- a dump made by this script only proves the oracle equals itself.
Usage: python tools/zig_golden/selftest_dump.py <dir> ; python -m pytest tests/test_zig_golden.py ; rm -r <dir>"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "zig")
os.makedirs(out, exist_ok=True)
orc = O.Oracle("128"); k = orc.keygen(7)
w = lambda name, a: np.ascontiguousarray(a).tofile(os.path.join(out, name))
w("bsk.bin", k.bsk); w("ksk.bin", k.ksk); w("secret.bin", np.concatenate([k.s0, k.s1]).astype(np.uint32))
ops, A, B = [], [], []
for p in range(3):
    for op in (O.NAND, O.AND, O.OR, O.XOR, O.NOR, O.XNOR):
        for row in range(4):
            ops.append(op); A.append((row >> 1) & 1); B.append(row & 1)
ops = np.array(ops, np.int32); A = np.array(A, np.uint8); B = np.array(B, np.uint8)
ca = orc.encrypt_bools(A, k, 1); cb = orc.encrypt_bools(B, k, 2)
go = orc.gate_batch(ops, ca, cb, k)
w("gate_ops.bin", ops); w("gate_a.bin", ca); w("gate_b.bin", cb); w("gate_out.bin", go)
w("gate_bits.bin", np.stack([A, B, orc.decrypt_bools(go, k)], axis=1).astype(np.uint8))
rin = orc.encrypt_bools(np.arange(8, dtype=np.uint8) & 1, k, 3)
rtr = orc.blind_rotate_batch(rin, k)
rl1 = np.stack([orc.sample_extract_index(t, 0) for t in rtr])
w("rot_in.bin", rin); w("rot_trlwe.bin", rtr); w("rot_lv1.bin", rl1); w("rot_lv0.bin", orc.keyswitch_batch(rl1, k))
open(os.path.join(out, "manifest.txt"), "w").write(f"{orc.n} 1024 {orc.L} {orc.bgbit} {orc.basebit} {orc.iks_t} {k.offset} {len(ops)} 8\n")
print("wrote", out, "(oracle self-test dump, NOT a pin)")
