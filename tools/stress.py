"""Randomised stress of the gate path: random batch sizes (all launch paths: cluster pair kernel, single-CTA latency kernel,
throughput kernel with tail launches), random per-item opcodes, decrypt-and-compare on every item, oracle equality on a
sample.  Usage: python tools/stress.py [rounds] [seed]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zig-tfhe_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tfhe_b200  # noqa: E402
from oracle import oracle as O  # noqa: E402
from conftest import TRUTH  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 60
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
orc = O.Oracle("128"); keys = orc.keygen(seed)
ctx = tfhe_b200.Context("128", devices=[0])
ctx.load_key(keys.bsk, keys.ksk, keys.offset)
pool = 4096
pa = rng.integers(0, 2, pool).astype(np.uint8); pb = rng.integers(0, 2, pool).astype(np.uint8)
ca = orc.encrypt_bools(pa, keys, 101); cb = orc.encrypt_bools(pb, keys, 102)
t0 = time.time()
checked = 0
for r in range(rounds):
    B = int(rng.choice([1, 2, 3, 37, 74, 75, 100, 148, 149, 300, 591, 592, 593, 700, 1185, 888, 889, 2048, int(rng.integers(1, 2500))]))
    idx = rng.integers(0, pool, B)
    ops = rng.integers(0, 10, B).astype(np.int32)
    ctx.set_tuning("latency_mode", int(rng.choice([0, 1, 1, 2])))
    ctx.set_tuning("kct", int(rng.choice([0, 0, 0, 1, 2, 3, 4, 6])))
    ctx.set_tuning("twt", int(rng.choice([0, 0, -1, 1])))              # tensor-memory twiddles: where they win / never / also at 4 and 5
    ctx.set_tuning("ks_tc", int(rng.choice([0, 0, -1, 1])))            # key switch: automatic / scalar kernel / tensor cores
    ctx.set_mode(tfhe_b200.MODE_EXACT if rng.integers(0, 6) == 0 else tfhe_b200.MODE_FAST)
    ctx.set_tuning("exact_kct", int(rng.choice([0, 0, 4, 6])))
    out = ctx.gate_batch(ops, ca[idx], cb[idx])
    want = np.array([TRUTH[int(ops[i])](int(pa[idx[i]]), int(pb[idx[i]])) for i in range(B)], np.uint8)
    got = orc.decrypt_bools(out, keys)
    assert (got == want).all(), f"round {r}: B={B}: {int((got != want).sum())} wrong bits"
    for i in rng.integers(0, B, min(B, 3)):
        ref = orc.gate(int(ops[i]), ca[idx[i]], cb[idx[i]], keys)
        assert (out[i] == ref).all(), f"round {r}: B={B}: item {i} differs from the oracle"
    checked += B
print(f"stress ok: {rounds} rounds, {checked} gates, {time.time() - t0:.1f} s")
ctx.close()
