"""ncu driver for the key-switch kernel alone (B lv1 samples, 128-bit set)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zig-tfhe_b200"))
import tfhe_b200  # noqa: E402
from tfhe_b200 import hostkeys as HK  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
params = tfhe_b200.PARAM_SETS["128"]
sk, ck = HK.gen_cloud_key(params, seed=1)
ctx = tfhe_b200.Context(params, devices=[0])
ctx.load_cloud_key(ck)
g = torch.Generator(device="cuda").manual_seed(0)
lv1 = torch.randint(-2**31, 2**31 - 1, (B, 1025), dtype=torch.int32, device="cuda", generator=g)
out = torch.empty((B, 701), dtype=torch.int32, device="cuda")
for _ in range(3):
    ctx.keyswitch_batch_device(0, lv1.data_ptr(), out.data_ptr(), B)
ctx.sync()
print("ok")
ctx.close()
