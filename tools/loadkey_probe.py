import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/zig-tfhe_b200")
import tfhe_b200
p = tfhe_b200.PARAM_SETS["128"]
rng = np.random.default_rng(0)
bsk = rng.standard_normal((p.n, 6, 2, 1024)); ksk = rng.integers(0, 2**32, (1024*9*4, 701), dtype=np.uint64).astype(np.uint32)
ctx = tfhe_b200.Context(p)
for i in range(5):
    t0 = time.perf_counter(); ctx.load_key(bsk, ksk, 0x82080000); ctx.sync(); print("load_key", i, f"{(time.perf_counter()-t0)*1e3:.0f} ms", flush=True)
tfhe_b200.key_file_write("/tmp/k.key", p, bsk, ksk, 1)
for i in range(4):
    t0 = time.perf_counter(); ctx.load_key_file("/tmp/k.key"); ctx.sync(); print("load_key_file", i, f"{(time.perf_counter()-t0)*1e3:.0f} ms", flush=True)
c2 = tfhe_b200.Context(p)
for i in range(3):
    t0 = time.perf_counter(); c2.load_key_file("/tmp/k.key"); c2.sync(); print("fresh ctx load_key_file", i, f"{(time.perf_counter()-t0)*1e3:.0f} ms", flush=True)
