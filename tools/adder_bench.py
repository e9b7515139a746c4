"""BASELINE config 3 as a measurement: 1,024 parallel 16-bit ripple-carry additions (81,920 gates, 33 levels),
host-level batching vs device-resident levels."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zig-tfhe_b200"))
import tfhe_b200  # noqa: E402
from tfhe_b200 import circuits, hostkeys as HK  # noqa: E402

W, B = 16, int(sys.argv[1]) if len(sys.argv) > 1 else 1024
params = tfhe_b200.PARAM_SETS["128"]
sk, ck = HK.gen_cloud_key(params, seed=1)
ctx = tfhe_b200.Context(params, devices=[0])
ctx.load_cloud_key(ck)
for kv in sys.argv[2:]:          # extra tuning keys, e.g. twt=-1 (round-1 kernels: four ciphertexts per CTA in the lanes)
    k, v = kv.split("=")
    ctx.set_tuning(k, int(v))
rng = np.random.default_rng(7)
x = rng.integers(0, 2**16, B); y = rng.integers(0, 2**16, B)
x[0], y[0] = 402, 304
enc = lambda bits: np.stack([HK.encrypt_bools(bits[i], params, sk, rng) for i in range(W)])
ca, cb = enc(circuits.to_bits(x, W)), enc(circuits.to_bits(y, W))
cin = HK.encrypt_bools(np.zeros(B, np.uint8), params, sk, rng)
t = lambda a: torch.from_numpy(a.view(np.int32)).cuda()
da, db, dc = t(ca), t(cb), t(cin)
circ = None
for name in ("host-level", "device-resident", "native circuit, 1 lane (CUDA graph)", "native circuit, 1 lane (eager levels)",
             "native circuit, 2 lanes (CUDA graph)", "native circuit, 4 lanes (CUDA graph)", "native circuit, 8 lanes (CUDA graph)"):
    if "lane" in name:
        lanes = int(name.split(",")[1].split()[0])
        if circ is None or lanes != cur_lanes:
            ctx.set_tuning("circuit_lanes", lanes)
            circ, cur_lanes = None, lanes
    for rep in range(2):
        t0 = time.perf_counter()
        if name == "host-level":
            sums, carry, gates, levels = circuits.ripple_carry_add(ctx, ca, cb, cin)
        elif name == "device-resident":
            s_d, c_d = circuits.ripple_carry_add_device(ctx, da, db, dc)
            sums, carry = s_d.cpu().numpy().view(np.uint32), c_d.cpu().numpy().view(np.uint32)
        else:
            ctx.set_tuning("circuit_graph", 1 if "graph" in name else 0)
            sums, carry, circ = circuits.ripple_carry_add_native(ctx, ca, cb, cin, circ)
        dt = time.perf_counter() - t0
    dec = np.stack([HK.decrypt_bools(sums[i], sk) for i in range(W)])
    total = circuits.from_bits(dec) + (HK.decrypt_bools(carry, sk).astype(np.uint64) << np.uint64(W))
    print(f"{name}: {B} additions x 80 gates in {dt * 1e3:.1f} ms = {80 * B / dt:.0f} gates/s, {dt / B * 1e3:.3f} ms per addition, correct={bool((total == x + y).all())}", flush=True)
ctx.close()
