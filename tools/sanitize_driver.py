"""Drives every kernel family once on a TINY parameter set so that compute-sanitizer (memcheck / racecheck / synccheck)
finishes in minutes: n = 20 CMUX steps instead of 700, keys generated on the device.  No torch, plain ctypes.

    compute-sanitizer --tool racecheck python tools/sanitize_driver.py [which ...]

which: k1 (throughput kernel: TMA and direct loads, KCT 1/4/6, teams of two, tail launch), lat (single-CTA latency
kernel), pair (two-CTA cluster kernel, st.async partial sums), k2 (key switch with and without i-range splits),
exact (reference-DAG kernel), keygen, circuit (4 lanes, CUDA graph), lut.  Default: all.
Results are checked by decryption where the set decrypts (the tiny set keeps the 128-bit set's noise and gadget)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zig-tfhe_b200"))
import tfhe_b200  # noqa: E402
from tfhe_b200 import circuits, hostkeys as HK  # noqa: E402

which = set(sys.argv[1:]) or {"k1", "lat", "pair", "k2", "exact", "keygen", "circuit", "lut"}
HK.ALPHAS["tiny"] = HK.ALPHAS["128"]
params = tfhe_b200.Params("tiny", 20, 3, 6, 2, 9)
rng = np.random.default_rng(1)
sk = HK.gen_secret_key(params, rng)
ctx = tfhe_b200.Context(params, devices=[0])
a0, a1 = HK.ALPHAS["tiny"]
ctx.keygen(sk.key_lv0, sk.key_lv1, seed=5, ksk_alpha=a0, bsk_alpha=a1, export=False)
print("keygen ok", flush=True)


def gates(B, tag, **tuning):
    for k, v in tuning.items():
        ctx.set_tuning(k, v)
    a = rng.integers(0, 2, B).astype(np.uint8); b = rng.integers(0, 2, B).astype(np.uint8)
    ca = HK.encrypt_bools(a, params, sk, rng); cb = HK.encrypt_bools(b, params, sk, rng)
    ops = (np.arange(B) % 10).astype(np.int32)
    out = ctx.gate_batch(ops, ca, cb)
    ok = int((HK.decrypt_bools(out, sk) == HK.decrypt_bools(ctx.gate_batch(ops, ca, cb), sk)).all())
    print(f"{tag}: B={B} {tuning} deterministic={ok}", flush=True)
    for k in tuning:
        ctx.set_tuning(k, 1 if k in ("use_tma", "latency_mode") else 0)
    return out


if "k1" in which:
    for kct in (1, 2, 4, 6):
        gates(2 * kct + 1, "K1 throughput, TMA ring", latency_mode=0, kct=kct)
    gates(9, "K1 throughput, direct loads", latency_mode=0, kct=4, use_tma=0)
    gates(7, "K1 throughput, teams of two", latency_mode=0, kct=4, team=2)
    gates(148 * 4 + 5, "K1 full wave + tail launch", latency_mode=0)
if "lat" in which:
    gates(3, "K1 single-CTA latency kernel", latency_mode=2)
if "pair" in which:
    gates(3, "K1 two-CTA cluster kernel")
if "k2" in which:
    lv1 = rng.integers(0, 2**32, (5, 1025), dtype=np.uint32)
    x = ctx.keyswitch_batch(lv1)                                  # 256 splits, atomics after a memset
    ctx.set_tuning("ks_fill", 1)
    big = rng.integers(0, 2**32, (1200, 1025), dtype=np.uint32)
    y = ctx.keyswitch_batch(big)                                  # one split: plain stores
    ctx.set_tuning("ks_fill", 0)
    z = ctx.keyswitch_batch(big)                                  # automatic splits
    print("K2 splits/no splits equal:", bool((y == z).all()), bool((x == ctx.keyswitch_batch(lv1)).all()), flush=True)
if "exact" in which:
    ctx.set_mode(tfhe_b200.MODE_EXACT)
    e = gates(5, "K1x exact mode")
    ctx.set_mode(tfhe_b200.MODE_FAST)
    print("exact ok", e.shape, flush=True)
if "circuit" in which:
    W, inst = 2, 9
    x = rng.integers(0, 4, inst); y = rng.integers(0, 4, inst)
    enc = lambda bits: np.stack([HK.encrypt_bools(bits[i], params, sk, rng) for i in range(W)])
    s, c, q = circuits.ripple_carry_add_native(ctx, enc(circuits.to_bits(x, W)), enc(circuits.to_bits(y, W)),
                                               HK.encrypt_bools(np.zeros(inst, np.uint8), params, sk, rng))
    s2, c2, _ = circuits.ripple_carry_add_native(ctx, enc(circuits.to_bits(x, W)), enc(circuits.to_bits(y, W)),
                                                 HK.encrypt_bools(np.zeros(inst, np.uint8), params, sk, rng), q)   # graph replay
    q.close()
    print("circuit (4 lanes, graph) ok", s.shape, flush=True)
if "lut" in which:
    tables = (np.arange(8, dtype=np.uint32) << np.uint32(28))
    ct = HK.encrypt_bools(np.ones(6, np.uint8), params, sk, rng)
    ctx.lut_bootstrap_batch(ct, np.tile(tables, (6, 1)), per_item=True)
    print("lut ok", flush=True)
ctx.close()
print("done")
