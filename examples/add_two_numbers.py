#!/usr/bin/env python
"""The reference's examples/add_two_numbers.zig on the B200 path: 402 + 304 = 706 under encryption.

The reference chains fullAdder gate by gate (80 bootstraps, ~3 s of CPU); here the same netlist is one circuit call
(tfhe_b200_circuit_run), and because instances are independent the same call adds as many pairs as you give it.

    python examples/add_two_numbers.py [instances]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zig-tfhe_b200"))
import tfhe_b200  # noqa: E402
from tfhe_b200 import circuits, hostkeys as HK  # noqa: E402

instances = int(sys.argv[1]) if len(sys.argv) > 1 else 1
W = 16
params = tfhe_b200.PARAM_SETS["128"]
rng = np.random.default_rng(2024)
print("=== TFHE Add Two Numbers Example (B200) ===")
sk = HK.gen_secret_key(params, rng)                       # key.SecretKey.new
ctx = tfhe_b200.Context(params)
t0 = time.perf_counter()
a_lv0, a_lv1 = HK.ALPHAS["128"]
ctx.keygen(sk.key_lv0, sk.key_lv1, seed=7, ksk_alpha=a_lv0, bsk_alpha=a_lv1, export=False)   # key.CloudKey.new, on the device
print(f"cloud key generated on the GPU in {(time.perf_counter() - t0) * 1e3:.1f} ms")

a = rng.integers(0, 2**W, instances); b = rng.integers(0, 2**W, instances)
a[0], b[0] = 402, 304                                     # add_two_numbers.zig:103-104
enc = lambda bits: np.stack([HK.encrypt_bools(bits[i], params, sk, rng) for i in range(W)])
ca, cb = enc(circuits.to_bits(a, W)), enc(circuits.to_bits(b, W))
cin = HK.encrypt_bools(np.zeros(instances, np.uint8), params, sk, rng)

t0 = time.perf_counter()
sums, carry, circ = circuits.ripple_carry_add_native(ctx, ca, cb, cin)
dt = time.perf_counter() - t0
bits = np.stack([HK.decrypt_bools(sums[i], sk) for i in range(W)])
total = circuits.from_bits(bits) + (HK.decrypt_bools(carry, sk).astype(np.uint64) << np.uint64(W))
print(f"A = {a[0]}, B = {b[0]}, decrypted sum = {total[0]} (expected {a[0] + b[0]})")
print(f"{instances} addition(s): {circ.n_gates} gates in {circ.levels} levels each, {dt * 1e3:.1f} ms "
      f"({circ.n_gates * instances / dt:.0f} bootstrapped gates/s)")
assert (total == a + b).all()
print("SUCCESS: homomorphic addition is correct")
ctx.close()
