"""K1 time versus number of CTA waves (does throughput hold once CTAs drift out of lock step?)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zig-tfhe_b200"))
import tfhe_b200  # noqa: E402
from tfhe_b200 import hostkeys as HK  # noqa: E402

params = tfhe_b200.PARAM_SETS["128"]
sk, ck = HK.gen_cloud_key(params, seed=1)
rng = np.random.default_rng(42)
Bmax = 148 * 6 * 74
ndist = int(os.environ.get("NDIST", "1024"))
a = rng.integers(0, 2, ndist).astype(np.uint8); b = rng.integers(0, 2, ndist).astype(np.uint8)
ca = HK.encrypt_bools(a, params, sk, rng); cb = HK.encrypt_bools(b, params, sk, rng)
CA = np.tile(ca, (Bmax // ndist + 1, 1))[:Bmax]; CB = np.tile(cb, (Bmax // ndist + 1, 1))[:Bmax]
ctx = tfhe_b200.Context(params, devices=[0])
ctx.load_cloud_key(ck)
ctx.set_tuning("timing", 1)
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    ctx.set_tuning(k, int(v))
for kct in [int(k) for k in os.environ.get("KCTS", "6,4").split(",")]:
    ctx.set_tuning("kct", kct)
    for waves in [int(w) for w in os.environ.get("WAVES", "1,2,4,8,16,37,74").split(",")]:
        B = 592 * waves if os.environ.get("FIXED_WAVE") else 148 * kct * waves
        ctx.gate_batch(tfhe_b200.NAND, CA[:B], CB[:B])
        ms = ctx.last_kernel_ms(0, 0)
        print(f"kct={kct} waves={waves:3d} B={B:6d} K1={ms:8.2f} ms  {ms / waves:6.2f} ms/wave  {B / ms * 1e3:9.0f} bootstraps/s", flush=True)
ctx.close()
