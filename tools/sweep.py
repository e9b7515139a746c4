"""BASELINE configs 4 and 5 as measurements (not bench lines): gate throughput at the 80/110/128-bit sets and the
UINT4 LUT bootstrap (exact mode, batch 32,768), device-resident, on every visible GPU of this process's context.

    python tools/sweep.py [gates_per_set] [uint4_batch]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zig-tfhe_b200"))
import tfhe_b200  # noqa: E402
from tfhe_b200 import hostkeys as HK  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
U = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
ndev = torch.cuda.device_count()
res = {"n_gpus": ndev, "results": []}
for name in ("80", "110", "128"):
    params = tfhe_b200.PARAM_SETS[name]
    rng = np.random.default_rng(42)
    sk = HK.gen_secret_key(params, rng)
    D = min(G, 1 << 16)                              # distinct ciphertext pairs, tiled up to G (host-side encryption is numpy)
    a = rng.integers(0, 2, D).astype(np.uint8); b = rng.integers(0, 2, D).astype(np.uint8)
    ca = HK.encrypt_bools(a, params, sk, rng); cb = HK.encrypt_bools(b, params, sk, rng)
    reps = (G + D - 1) // D
    a = np.tile(a, reps)[:G]; b = np.tile(b, reps)[:G]
    ca = np.tile(ca, (reps, 1))[:G]; cb = np.tile(cb, (reps, 1))[:G]
    ctx = tfhe_b200.Context(params, devices=list(range(ndev)))
    ctx.keygen(sk.key_lv0, sk.key_lv1, seed=1, ksk_alpha=HK.ALPHAS[name][0], bsk_alpha=HK.ALPHAS[name][1], export=False)   # every device generates the same key
    ctx.gate_batch(tfhe_b200.NAND, ca, cb)          # warm-up at full size: scratch buffers allocated, clocks up
    t0 = time.perf_counter()
    out = ctx.gate_batch(tfhe_b200.NAND, ca, cb)
    dt = time.perf_counter() - t0
    ok = bool((HK.decrypt_bools(out, sk) == 1 - (a & b)).all())
    res["results"].append({"config": f"{G} NAND gates, SECURITY_{name}_BIT, host buffers (e2e), {ndev} GPU(s) in one context",
                           "gates_per_s": G / dt, "seconds": dt, "all_bits_correct": ok})
    print(res["results"][-1], flush=True)
    ctx.close()
    del ca, cb, out
# config 4: UINT4 LUT bootstrap, exact mode
params = tfhe_b200.PARAM_SETS["uint4"]
sk, ck = HK.gen_cloud_key(params, seed=1)
rng = np.random.default_rng(7)
ct = HK.tlwe_encrypt_f64(rng.integers(0, 16, U) / 32.0, HK.ALPHAS["uint4"][0], sk.key_lv0, rng)
tv = np.zeros((2, 1024), np.uint32)
tv[1] = (np.arange(1024) // 64).astype(np.uint32) << np.uint32(27)      # identity-like staircase test vector
for mode, label in ((tfhe_b200.MODE_EXACT, "exact"), (tfhe_b200.MODE_FAST, "fast")):
    ctx = tfhe_b200.Context(params, devices=list(range(ndev)))
    ctx.load_cloud_key(ck)
    ctx.set_mode(mode)
    ctx.bootstrap_batch(ct, tv)                     # warm-up at full size
    t0 = time.perf_counter()
    out = ctx.bootstrap_batch(ct, tv)
    dt = time.perf_counter() - t0
    res["results"].append({"config": f"{U} LUT bootstraps, SECURITY_UINT4, {label} mode, {ndev} GPU(s)", "bootstraps_per_s": U / dt, "seconds": dt})
    print(res["results"][-1], flush=True)
    ctx.close()
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"sweep_{ndev}gpu.json"), "w"), indent=1)
