"""Tiny driver for ncu captures: a few waves of NAND gates through the device path (B and kct from argv)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zig-tfhe_b200"))
import tfhe_b200  # noqa: E402
from tfhe_b200 import hostkeys as HK  # noqa: E402

kct = int(sys.argv[1]) if len(sys.argv) > 1 else 0
B = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 6 * 2
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
params = tfhe_b200.PARAM_SETS["128"]
sk, ck = HK.gen_cloud_key(params, seed=1)
rng = np.random.default_rng(42)
a = rng.integers(0, 2, B).astype(np.uint8); b = rng.integers(0, 2, B).astype(np.uint8)
ca = HK.encrypt_bools(a, params, sk, rng); cb = HK.encrypt_bools(b, params, sk, rng)
ctx = tfhe_b200.Context(params, devices=[0])
ctx.load_cloud_key(ck)
if kct:
    ctx.set_tuning("kct", kct)
for kv in sys.argv[4:]:          # extra tuning keys: pipeline=0 team=2 ...
    k, v = kv.split("=")
    ctx.set_tuning(k, int(v))
ctx.set_tuning("timing", 1)
for _ in range(reps):
    out = ctx.gate_batch(tfhe_b200.NAND, ca, cb)
    print("K1 ms", ctx.last_kernel_ms(0, 0), "K2 ms", ctx.last_kernel_ms(0, 1), "->", B / ctx.last_kernel_ms(0, 0) * 1e3, "bootstraps/s")
if any(kv.startswith("diag=") and kv != "diag=0" for kv in sys.argv[4:]):
    print("diagnostic run: results are wrong on purpose")
else:
    assert (HK.decrypt_bools(out, sk) == 1 - (a & b)).all()
    print("ok")
ctx.close()
