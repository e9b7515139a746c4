"""Step-by-step bring-up probe for a fresh B200 box: each stage prints PASS/FAIL on its own line so one
failure does not hide the rest.  Not part of the test suite (tests/ is)."""
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
    stage("fp64 peak TFLOP/s", lambda: ctx.measure_fp64_tflops(0))
    stage("load key", lambda: ctx.load_key(keys.bsk, keys.ksk, keys.offset))
    rng = np.random.default_rng(0)
    bits_a = rng.integers(0, 2, 64).astype(np.uint8); bits_b = rng.integers(0, 2, 64).astype(np.uint8)
    ca = orc.encrypt_bools(bits_a, keys, 1); cb = orc.encrypt_bools(bits_b, keys, 2)
    lin = np.stack([orc.gate_linear(O.NAND, ca[i], cb[i]) for i in range(8)])
    ref_tr = orc.blind_rotate_batch(lin, keys)

    def ks():
        lv1 = rng.integers(0, 2**32, (33, 1025), dtype=np.uint32)
        assert (ctx.keyswitch_batch(lv1) == orc.keyswitch_batch(lv1, keys)).all()
    stage("keyswitch bit-exact", ks)

    def br(tma, kct, B):
        def f():
            ctx.set_tuning("use_tma", tma); ctx.set_tuning("kct", kct)
            got = ctx.blind_rotate_batch(lin[:B])
            nd = int((got != ref_tr[:B]).sum())
            assert nd == 0, f"{nd} coefficients differ"
        return f
    stage("blind rotate direct-load kct=1 B=1", br(0, 1, 1))
    stage("blind rotate direct-load kct=2 B=3", br(0, 2, 3))
    stage("blind rotate TMA kct=1 B=1", br(1, 1, 1))
    stage("blind rotate TMA kct=2 B=3", br(1, 2, 3))
    stage("blind rotate TMA kct=6 B=8", br(1, 6, 8))
    stage("blind rotate direct kct=6 B=8", br(0, 6, 8))
    ctx.set_tuning("use_tma", 1); ctx.set_tuning("kct", 0)

    def gates():
        out = ctx.gate_batch(O.NAND, ca, cb)
        assert (orc.decrypt_bools(out, keys) == 1 - (bits_a & bits_b)).all()
        assert (out[:8] == orc.gate_batch(O.NAND, ca[:8], cb[:8], keys)).all()
    stage("64 NAND gates decrypt + bit-exact", gates)

    # quick throughput sweep (host API, includes copies) -- tuning signal only
    big = 148 * 6 * 4
    A = np.tile(ca, (big // 64 + 1, 1))[:big]; Bm = np.tile(cb, (big // 64 + 1, 1))[:big]
    ctx.set_tuning("timing", 1)
    for tma in (1, 0):
        for kct in (6, 5, 4, 3, 2):
            def run(tma=tma, kct=kct):
                ctx.set_tuning("use_tma", tma); ctx.set_tuning("kct", kct)
                nb = 148 * kct * 2
                ctx.gate_batch(O.NAND, A[:nb], Bm[:nb])
                ctx.gate_batch(O.NAND, A[:nb], Bm[:nb])
                k1, k2 = ctx.last_kernel_ms(0, 0), ctx.last_kernel_ms(0, 1)
                return f"B={nb} K1={k1:.2f}ms K2={k2:.2f}ms -> {nb / (k1 * 1e-3):.0f} bootstraps/s (K1 only)"
            stage(f"throughput tma={tma} kct={kct}", run)
    ctx.close()


if __name__ == "__main__":
    main()
