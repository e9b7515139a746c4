"""Cloud-key generation: device (tfhe_b200_keygen) vs the numpy host mirror and the C++ oracle, 128-bit set."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zig-tfhe_b200"))
import tfhe_b200  # noqa: E402
from tfhe_b200 import hostkeys as HK  # noqa: E402
from oracle import oracle as O  # noqa: E402

params = tfhe_b200.PARAM_SETS["128"]
t0 = time.perf_counter(); sk, ck_host = HK.gen_cloud_key(params, seed=1); t_host = time.perf_counter() - t0
t0 = time.perf_counter(); keys = O.Oracle("128").keygen(seed=1); t_orc = time.perf_counter() - t0
ctx = tfhe_b200.Context(params, devices=[0])
a0, a1 = HK.ALPHAS["128"]
for export in (False, True, False):
    t0 = time.perf_counter()
    ck = ctx.keygen(keys.s0, keys.s1, seed=7, ksk_alpha=a0, bsk_alpha=a1, export=export)
    ctx.sync()
    dt = time.perf_counter() - t0
    print(f"device keygen (export to host = {export}): {dt * 1e3:.1f} ms", flush=True)
# flat cloud-key file: write, then a fresh context loads it (mmap + checksum + upload + re-layout)
import tempfile  # noqa: E402
ck = ctx.keygen(keys.s0, keys.s1, seed=7, ksk_alpha=a0, bsk_alpha=a1, export=True)
with tempfile.TemporaryDirectory(dir=os.environ.get("KEYFILE_DIR")) as tmp:
    path = os.path.join(tmp, "cloud128.key")
    t0 = time.perf_counter(); ck.save(path, "128"); t_save = time.perf_counter() - t0
    ctx2 = tfhe_b200.Context(params, devices=[0])
    for _ in range(2):
        t0 = time.perf_counter(); ctx2.load_key_file(path); ctx2.sync(); t_load = time.perf_counter() - t0
        print(f"key file ({os.path.getsize(path) / 1e6:.1f} MB): save {t_save * 1e3:.0f} ms, load_key_file {t_load * 1e3:.0f} ms", flush=True)
    t0 = time.perf_counter(); ctx2.load_key(ck.bootstrapping_key, ck.key_switching_key, ck.decomposition_offset); ctx2.sync()
    print(f"load_key from host arrays: {(time.perf_counter() - t0) * 1e3:.0f} ms")
    ctx2.close()
print(f"numpy host mirror: {t_host:.2f} s; C++ oracle (all host threads): {t_orc:.2f} s; reference (key.zig:240 comment): ~30 s")
ctx.close()
