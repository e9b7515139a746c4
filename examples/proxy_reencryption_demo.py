#!/usr/bin/env python
"""The reference's examples/proxy_reencryption_demo.zig on the B200 path: Alice's ciphertexts are re-encrypted for Bob
by a proxy that never sees a plaintext (proxy_reenc.reencryptTLWELv0, src/proxy_reenc.zig:267-306), as ONE batched call
(tfhe_b200_reencrypt_batch = the key-switch kernel with source dimension n).

    python examples/proxy_reencryption_demo.py [count]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zig-tfhe_b200"))
import tfhe_b200  # noqa: E402
from tfhe_b200 import hostkeys as HK  # noqa: E402

count = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
params = tfhe_b200.PARAM_SETS["128"]
rng = np.random.default_rng(11)
print("=== Proxy Reencryption Demo (B200) ===")
alice = HK.gen_secret_key(params, rng); bob = HK.gen_secret_key(params, rng)
rk = HK.gen_reencryption_key(params, alice.key_lv0, bob.key_lv0, rng)       # Alice -> Bob (symmetric mode)
ctx = tfhe_b200.Context(params)
a0, a1 = HK.ALPHAS["128"]
ctx.keygen(alice.key_lv0, alice.key_lv1, seed=3, ksk_alpha=a0, bsk_alpha=a1, export=False)
ctx.load_reencryption_key(rk)

bits = rng.integers(0, 2, count).astype(np.uint8)
ct_alice = HK.encrypt_bools(bits, params, alice, rng)
t0 = time.perf_counter()
ct_bob = ctx.reencrypt_batch(ct_alice)
dt = time.perf_counter() - t0
ok_bob = bool((HK.decrypt_bools(ct_bob, bob) == bits).all())
print(f"{count} ciphertexts re-encrypted in {dt * 1e3:.2f} ms ({count / dt:.0f} per second); Bob decrypts all correctly: {ok_bob}")
assert ok_bob
# the re-encrypted ciphertexts are ordinary TLWE samples under Bob's key: Bob's server can keep computing on them
ctx_bob = tfhe_b200.Context(params)
ctx_bob.keygen(bob.key_lv0, bob.key_lv1, seed=4, ksk_alpha=a0, bsk_alpha=a1, export=False)
half = count // 2
out = ctx_bob.gate_batch(tfhe_b200.XOR, ct_bob[:half], ct_bob[half:2 * half])
assert (HK.decrypt_bools(out, bob) == (bits[:half] ^ bits[half:2 * half])).all()
print(f"XOR of {half} re-encrypted pairs under Bob's cloud key: correct")
print("SUCCESS")
ctx.close(); ctx_bob.close()
