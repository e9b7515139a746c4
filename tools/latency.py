"""Per-bootstrap latency of small batches (device resident, CUDA events): latency mode vs throughput mode."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zig-tfhe_b200"))
import tfhe_b200  # noqa: E402
from tfhe_b200 import hostkeys as HK  # noqa: E402

params = tfhe_b200.PARAM_SETS["128"]
sk, ck = HK.gen_cloud_key(params, seed=1)
ctx = tfhe_b200.Context(params, devices=[0])
ctx.load_cloud_key(ck)
ctx.set_tuning("timing", 1)
rng = np.random.default_rng(0)
Bmax = 2048
a = rng.integers(0, 2, Bmax).astype(np.uint8); b = rng.integers(0, 2, Bmax).astype(np.uint8)
d_a = torch.from_numpy(HK.encrypt_bools(a, params, sk, rng).view(np.int32)).cuda()
d_b = torch.from_numpy(HK.encrypt_bools(b, params, sk, rng).view(np.int32)).cuda()
d_o = torch.empty_like(d_a)
stream = torch.cuda.ExternalStream(ctx.stream(0))
for mode in (1, 2, 0):
    ctx.set_tuning("latency_mode", mode)
    for B in (1, 16, 74, 148, 296, 592, 2048):
        ts = []
        for _ in range(12):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.gate_batch_device(0, tfhe_b200.NAND, None, d_a.data_ptr(), d_b.data_ptr(), d_o.data_ptr(), B)
            e1.record(stream)
            ctx.sync()
            ts.append(e0.elapsed_time(e1))
        ok = bool((HK.decrypt_bools(d_o[:B].cpu().numpy().view(np.uint32), sk) == 1 - (a[:B] & b[:B])).all())
        print(f"latency_mode={mode} B={B:5d} p50={np.median(ts[2:]):7.3f} ms  K1={ctx.last_kernel_ms(0, 0):6.3f} K2={ctx.last_kernel_ms(0, 1):6.3f}  ({B / np.median(ts[2:]) * 1e3:9.0f} gates/s) ok={ok}", flush=True)
ctx.close()
