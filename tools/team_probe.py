"""Probe of the K1 variants behind the tuning keys kct / team: bit-exact parity against the oracle on ragged batch
sizes, cross-check against the default kernel on every coefficient, then a K1-only throughput comparison.
Not part of the test suite."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zig-tfhe_b200"))

import tfhe_b200  # noqa: E402
from oracle import oracle as O  # noqa: E402


def stage(name, fn):
    t = time.time()
    try:
        r = fn()
        print(f"[PASS] {name} ({time.time() - t:.2f}s) {r if r is not None else ''}", flush=True)
        return True
    except Exception as e:
        print(f"[FAIL] {name}: {e}", flush=True)
        traceback.print_exc()
        return False


def main():
    orc = O.Oracle("128")
    keys = orc.keygen(1)
    ctx = tfhe_b200.Context("128", devices=[0])
    ctx.load_key(keys.bsk, keys.ksk, keys.offset)
    ctx.set_tuning("latency_mode", 0)
    rng = np.random.default_rng(0)
    nref = 11
    bits_a = rng.integers(0, 2, 64).astype(np.uint8); bits_b = rng.integers(0, 2, 64).astype(np.uint8)
    ca = orc.encrypt_bools(bits_a, keys, 1); cb = orc.encrypt_bools(bits_b, keys, 2)
    lin = np.stack([orc.gate_linear(O.NAND, ca[i], cb[i]) for i in range(nref)])
    ref_tr = orc.blind_rotate_batch(lin, keys)

    def br(team, kct, B):
        def f():
            ctx.set_tuning("team", team); ctx.set_tuning("kct", kct)
            got = ctx.blind_rotate_batch(lin[:B])
            nd = int((got != ref_tr[:B]).sum())
            assert nd == 0, f"{nd} coefficients differ"
        return f
    for (team, kct, B) in [(1, 1, 1), (1, 2, 3), (1, 3, 5), (2, 2, 1), (2, 2, 3), (2, 4, 2), (2, 4, 7), (2, 4, 11), (1, 4, 11), (2, 6, 1), (2, 6, 5), (2, 6, 11)]:
        stage(f"blind rotate team={team} kct={kct} B={B} bit-exact vs oracle", br(team, kct, B))

    # larger ragged batch: every variant vs the default kernel, every coefficient
    big = 148 * 4 * 2 + 3
    A = np.tile(ca, (big // 64 + 1, 1))[:big]; Bm = np.tile(cb, (big // 64 + 1, 1))[:big]
    linb = np.stack([orc.gate_linear(O.XOR, A[i], Bm[i]) for i in range(big)])

    def cross():
        ctx.set_tuning("kct", 4); ctx.set_tuning("team", 1)
        r0 = ctx.blind_rotate_batch(linb)
        for (pipe, team, kct) in ((0, 2, 4), (0, 2, 6), (0, 1, 6), (0, 1, 3)):
            ctx.set_tuning("team", team); ctx.set_tuning("kct", kct)
            r1 = ctx.blind_rotate_batch(linb)
            assert (r0 == r1).all(), f"team={team}: {(r0 != r1).sum()} coefficients differ from the default kernel"
    stage(f"variants == default kernel on B={big}", cross)

    ctx.set_tuning("timing", 1)
    nb = 148 * 6 * 4
    A = np.tile(ca, (nb // 64 + 1, 1))[:nb]; Bm = np.tile(cb, (nb // 64 + 1, 1))[:nb]
    for (pipe, team, kct) in [(0, 1, 4), (0, 2, 4), (0, 2, 6), (0, 1, 6), (0, 1, 2), (0, 2, 2)]:
        def run(pipe=pipe, team=team, kct=kct):
            ctx.set_tuning("team", team); ctx.set_tuning("kct", kct)
            n = 148 * kct * 4
            ctx.gate_batch(O.NAND, A[:n], Bm[:n])
            ctx.gate_batch(O.NAND, A[:n], Bm[:n])
            k1, k2 = ctx.last_kernel_ms(0, 0), ctx.last_kernel_ms(0, 1)
            return f"B={n} K1={k1:.2f}ms K2={k2:.2f}ms -> {n / (k1 * 1e-3):.0f} bootstraps/s (K1 only)"
        stage(f"throughput team={team} kct={kct}", run)
    ctx.close()


if __name__ == "__main__":
    main()
