"""Key-switch kernel (K2) tile sweep at the BASELINE batch (65,536 lv1 samples, 128-bit set)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zig-tfhe_b200"))
import tfhe_b200  # noqa: E402
from tfhe_b200 import hostkeys as HK  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
params = tfhe_b200.PARAM_SETS["128"]
sk, ck = HK.gen_cloud_key(params, seed=1)
ctx = tfhe_b200.Context(params, devices=[0])
ctx.load_cloud_key(ck)
g = torch.Generator(device="cuda").manual_seed(0)
lv1 = torch.randint(-2**31, 2**31 - 1, (B, 1025), dtype=torch.int32, device="cuda", generator=g)
out = torch.empty((B, 701), dtype=torch.int32, device="cuda")
stream = torch.cuda.ExternalStream(ctx.stream(0))
ref = None
FILLS = [int(x) for x in os.environ.get("FILLS", "0").split(",")]     # CTAs per SM the i-range split aims for (0 = automatic)
TILES = [tuple(int(y) for y in x.split("x")) for x in os.environ.get("TILES", "8x1,8x2,4x2,4x1,16x1").split(",")]
ROTS = [int(x) for x in os.environ.get("ROTS", "0").split(",")]
for tile, vec, fill, rot in [(t, v, f, r) for (t, v) in TILES for f in FILLS for r in ROTS]:
    ctx.set_tuning("ks_rot", rot)
    ctx.set_tuning("ks_tile", tile)
    ctx.set_tuning("ks_vec", vec)
    ctx.set_tuning("ks_fill", fill)
    for _ in range(2):
        ctx.keyswitch_batch_device(0, lv1.data_ptr(), out.data_ptr(), B)
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3):
        ctx.keyswitch_batch_device(0, lv1.data_ptr(), out.data_ptr(), B)
    e1.record(stream)
    ctx.sync()
    ms = e0.elapsed_time(e1) / 3
    o = out.cpu().numpy()
    if ref is None:
        ref = o.copy()
    print(f"tile={tile:2d} vec={vec} fill={fill:2d} rot={rot} B={B} K2={ms:.2f} ms  ({B / ms * 1e3:.0f} keyswitch/s)  same_bits={bool((o == ref).all())}", flush=True)
ctx.close()
