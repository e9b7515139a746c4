"""Small workload touching every kernel once (device keygen, cluster pair kernel, single-CTA latency kernel, throughput
kernel with a tail launch, key switch, circuit executor, LUT generator), each result checked by decryption."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zig-tfhe_b200"))
import tfhe_b200  # noqa: E402
from tfhe_b200 import circuits, hostkeys as HK  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
params = tfhe_b200.PARAM_SETS["128"]
rng = np.random.default_rng(1)
sk = HK.gen_secret_key(params, rng)
ctx = tfhe_b200.Context(params, devices=[0])
a0, a1 = HK.ALPHAS["128"]
ck = ctx.keygen(sk.key_lv0, sk.key_lv1, seed=5, ksk_alpha=a0, bsk_alpha=a1, export=False)
sizes = {"pair": [3], "latency": [80], "throughput": [151], "all": [3, 80, 151]}[which]
for B in sizes:
    if which == "throughput" or (which == "all" and B == 151):
        ctx.set_tuning("latency_mode", 0)
    bits_a = rng.integers(0, 2, B).astype(np.uint8); bits_b = rng.integers(0, 2, B).astype(np.uint8)
    ca = HK.encrypt_bools(bits_a, params, sk, rng); cb = HK.encrypt_bools(bits_b, params, sk, rng)
    out = ctx.gate_batch(tfhe_b200.NAND, ca, cb)
    assert (HK.decrypt_bools(out, sk) == 1 - (bits_a & bits_b)).all(), B
    print("gate batch", B, "ok", flush=True)
if which == "all":
    ctx.set_tuning("latency_mode", 1)
    W, inst = 2, 3
    x = rng.integers(0, 4, inst); y = rng.integers(0, 4, inst)
    enc = lambda bits: np.stack([HK.encrypt_bools(bits[i], params, sk, rng) for i in range(W)])
    s, c, q = circuits.ripple_carry_add_native(ctx, enc(circuits.to_bits(x, W)), enc(circuits.to_bits(y, W)),
                                               HK.encrypt_bools(np.zeros(inst, np.uint8), params, sk, rng))
    dec = np.stack([HK.decrypt_bools(s[i], sk) for i in range(W)])
    assert (circuits.from_bits(dec) + (HK.decrypt_bools(c, sk).astype(np.uint64) << np.uint64(W)) == x + y).all()
    q.close()
    print("circuit ok", flush=True)
    tv = ctx.lut_generate(np.arange(4, dtype=np.uint32) << np.uint32(29))
    print("lut ok", int(tv[1, 0]), flush=True)
ctx.close()
print("done")
