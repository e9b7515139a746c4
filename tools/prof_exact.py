"""Tiny driver for ncu captures of the exact-mode kernel: UINT4 LUT bootstraps (B from argv) through the host path.
With a parameter-set name as second argument (e.g. uint8): blind rotations only (no key-switching key, 1.8 GB at UINT7/8), plus tuning keys k=v."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zig-tfhe_b200"))
import tfhe_b200  # noqa: E402
from tfhe_b200 import hostkeys as HK  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 4 * 2
name = sys.argv[2] if len(sys.argv) > 2 else "uint4"
params = tfhe_b200.PARAM_SETS[name]
sk, ck = HK.gen_cloud_key(params, seed=1, with_ksk=(name == "uint4"))
rng = np.random.default_rng(7)
ct = HK.tlwe_encrypt_f64(rng.integers(0, 16, B) / 32.0, HK.ALPHAS[name][0], sk.key_lv0, rng)
tv = np.zeros((2, 1024), np.uint32)
tv[1] = (np.arange(1024) // 64).astype(np.uint32) << np.uint32(27)
ctx = tfhe_b200.Context(params, devices=[0])
ctx.load_cloud_key(ck)
ctx.set_mode(tfhe_b200.MODE_EXACT)
ctx.set_tuning("timing", 1)
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    if k == "mode":          # mode=fast: the production kernel on the same inputs
        ctx.set_mode(tfhe_b200.MODE_FAST if v == "fast" else tfhe_b200.MODE_EXACT)
    else:
        ctx.set_tuning(k, int(v))
for _ in range(2):
    out = ctx.bootstrap_batch(ct, tv) if name == "uint4" else ctx.blind_rotate_batch(ct, tv)
    print("K1x ms", ctx.last_kernel_ms(0, 0), "K2 ms", ctx.last_kernel_ms(0, 1), "->", B / ctx.last_kernel_ms(0, 0) * 1e3, "bootstraps/s")
print("ok")
ctx.close()
