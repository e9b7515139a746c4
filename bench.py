#!/usr/bin/env python
"""bench.py -- bootstrapped gates/s of the zig-tfhe gate-bootstrapping hot path on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W]             (N = 1)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...                            (host-CPU arm)

Default workload (BASELINE.json configs[1]): 65,536 independent AND/XOR gates (half each) at
SECURITY_128_BIT per GPU ("weak" scaling: every rank owns its own 65,536-gate shard; keys are
generated once on rank 0 and broadcast with one NCCL broadcast at key load; there is no collective
on the hot path).  One step = one pass of the whole path (gate linear part -> blind rotation ->
sample extract -> key switch) over the batch.

The other BASELINE configs from the same build (the default line is unchanged by these flags):
  --params {80,110,128}            parameter set of the gate workload (configs[4])
  --params uint4 [--mode exact|fast]  programmable LUT bootstraps at SECURITY_UINT4, 32,768 per GPU (configs[3]);
                                   exact mode (the reference's own transform DAG) is the one that equals the oracle there
  --total N --scaling strong       N units in total, a contiguous shard of N / world per rank (configs[4]: --total 1048576)

  value : gates/s with inputs already resident in HBM (device entry points, CUDA events on the
          library's stream, max over ranks).
  e2e   : gates/s through the host-buffer C-ABI call (tfhe_b200_gate_batch_ops) from pinned host
          memory, H2D and D2H copies inside the timed region.
  roofline : blind-rotation kernel against the FP64-FMA roofline (north_star: no tensor cores; the
          DFMA peak is measured live by the library's microbenchmark because MEASURED_PEAKS.json
          has no FP64 figure) plus the HBM side (key bytes per wave).
  cpu_baseline : the CPU oracle (port of the reference) on this box's host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "zig-tfhe_b200"))

PARAMS = "128"
BATCH = 65536
LUT_BATCH = 32768
UNIT = "gates/s"
AND, XOR = 2, 3
LUT_MODULUS = 16


def metric_name(pname, mode):
    if pname.startswith("uint"):
        return f"LUT bootstraps/sec (whole box, {pname.upper()}, {mode} mode)"
    return f"bootstrapped gates/sec (whole box, {pname}-bit)"


def work_per_bootstrap(p):
    """algorithmic work per bootstrap, SURVEY.md section 8d: n x ((2L+2) transforms x 26,112 flop + 2L x 2 x 512 complex MACs x 8 flop);
    key bytes as in SURVEY.md appendix B"""
    flop = p.n * ((2 * p.L + 2) * 26112 + 2 * p.L * 2 * 512 * 8)
    bsk = p.n * 2 * p.L * 2 * 1024 * 8
    ksk = 1024 * p.iks_t * (1 << p.basebit) * (p.n + 1) * 4
    return flop, bsk, ksk


def synth_inputs(params, sk, batch, seed):
    """gate workload: `batch` AND/XOR gates (half each) on encryptions of uniform random bits.  Above 65,536 the distinct
    ciphertext pairs are tiled (host-side encryption is numpy; the device work does not depend on the values)."""
    from tfhe_b200 import hostkeys as HK
    rng = np.random.default_rng(seed)
    D = min(batch, 65536)
    a = rng.integers(0, 2, D).astype(np.uint8)
    b = rng.integers(0, 2, D).astype(np.uint8)
    ca = HK.encrypt_bools(a, params, sk, rng)
    cb = HK.encrypt_bools(b, params, sk, rng)
    if batch > D:
        reps = -(-batch // D)
        a, b = np.tile(a, reps)[:batch], np.tile(b, reps)[:batch]
        ca, cb = np.tile(ca, (reps, 1))[:batch], np.tile(cb, (reps, 1))[:batch]
    ops = np.where(np.arange(batch) < batch // 2, AND, XOR).astype(np.int32)
    truth = np.where(ops == AND, a & b, a ^ b).astype(np.uint8)
    return ca, cb, ops, truth


def synth_lut_inputs(params, sk, batch, seed):
    """LUT workload (BASELINE configs[3]): encryptions of uniform messages in [0, 16) (tlwe.encryptLweMessage, src/tlwe.zig:72-90)
    and the test vector of f(x) = x^2 mod 16 (lut/generator.zig:85-135)"""
    from tfhe_b200 import hostkeys as HK
    from tfhe_b200 import lut as LUT
    rng = np.random.default_rng(seed)
    msgs = rng.integers(0, LUT_MODULUS, batch)
    ct = HK.tlwe_encrypt_f64(msgs / (2.0 * LUT_MODULUS), HK.ALPHAS[params.name][0], sk.key_lv0, rng)
    tv = LUT.Generator(LUT_MODULUS).generate_lookup_table(lambda x: (x * x) % LUT_MODULUS).poly
    return ct, np.ascontiguousarray(tv, dtype=np.uint32)


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = sorted(sm)[len(sm) // 4:] if len(sm) > 4 else sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def oracle_for(pname, sk, ck):
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    orc = O.Oracle(pname)
    keys = O.Keys(sk.key_lv0, sk.key_lv1, ck.bootstrapping_key, ck.key_switching_key, ck.decomposition_offset, ck.blind_rotate_testvec)
    return O, orc, keys


def cpu_run(orc, keys, work, lo, hi, nthreads):
    """one pass of the CPU oracle (port of the reference path) over items [lo, hi) of the workload"""
    if work["kind"] == "lut":
        return orc.bootstrap_batch(work["ct"][lo:hi], keys, work["tv"], nthreads=nthreads)
    return orc.gate_batch(work["ops"][lo:hi], work["ca"][lo:hi], work["cb"][lo:hi], keys, nthreads=nthreads)


def cpu_reference_rate(pname, keys_tuple, work, seconds_target=12.0):
    """times the CPU oracle with all host threads on a bounded sample, and one thread on a few items (ms per bootstrap,
    the figure the reference's CHANGELOG quotes)"""
    O, orc, keys = oracle_for(pname, *keys_tuple)
    cores = O.hardware_threads()
    n_items = work["n"]
    probe = min(n_items, 2 * cores)
    t0 = time.perf_counter()
    cpu_run(orc, keys, work, 0, probe, cores)
    per_round = (time.perf_counter() - t0) / 2.0          # seconds for `cores` items
    sample = int(min(n_items, max(4 * cores, cores * max(1, int(seconds_target / max(per_round, 1e-3))))))
    t0 = time.perf_counter()
    out = cpu_run(orc, keys, work, 0, sample, cores)
    dt = time.perf_counter() - t0
    k1 = min(n_items, 8)
    t0 = time.perf_counter()
    cpu_run(orc, keys, work, 0, k1, 1)
    single_ms = (time.perf_counter() - t0) * 1e3 / k1
    return sample / dt, cores, sample, out, single_ms


def make_work(pname, params, sk, batch, seed):
    if pname.startswith("uint"):
        ct, tv = synth_lut_inputs(params, sk, batch, seed)
        return {"kind": "lut", "n": batch, "ct": ct, "tv": tv}
    ca, cb, ops, truth = synth_inputs(params, sk, batch, seed)
    return {"kind": "gates", "n": batch, "ca": ca, "cb": cb, "ops": ops, "truth": truth}


def workload_text(pname, B, mode, scaling, total, world):
    per = f"{B} per GPU" if scaling == "weak" else f"{total} in total, a contiguous shard of {B} per GPU"
    if pname.startswith("uint"):
        return f"programmable LUT bootstraps (x^2 mod 16, shared test vector) at SECURITY_{pname.upper()}, {mode} mode, {per}"
    return f"independent AND/XOR gates (half/half) at SECURITY_{pname}_BIT, {per}, seeded keys and ciphertexts"


def run_reference(args):
    """--impl reference: the reference's own algorithm on the host cores (oracle port; Zig cannot be built here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import tfhe_b200
    from tfhe_b200 import hostkeys as HK
    pname = args.params
    params = tfhe_b200.PARAM_SETS[pname]
    sk, ck = HK.gen_cloud_key(params, seed=1)
    sample_max = 4096
    work = make_work(pname, params, sk, sample_max, seed=42)
    rates = []
    O, orc, keys = oracle_for(pname, sk, ck)
    cores = O.hardware_threads()
    total = args.warmup + args.steps
    budget = 150.0
    t_probe = time.perf_counter()
    cpu_run(orc, keys, work, 0, cores, cores)
    per_round = time.perf_counter() - t_probe
    rounds = max(1, min(16, int(budget / total / max(per_round, 1e-3))))
    sample = min(sample_max, rounds * cores)
    ok = True
    for step in range(total):
        t0 = time.perf_counter()
        out = cpu_run(orc, keys, work, 0, sample, cores)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            rates.append(sample / dt)
        if work["kind"] == "gates":
            ok = ok and bool((HK.decrypt_bools(out, sk) == work["truth"][:sample]).all())
    t0 = time.perf_counter()
    cpu_run(orc, keys, work, 0, 4, 1)
    single_ms = (time.perf_counter() - t0) * 1e3 / 4
    value = float(np.mean(rates))
    B = args.batch or (LUT_BATCH if work["kind"] == "lut" else BATCH)
    unit = UNIT if work["kind"] == "gates" else "bootstraps/s"
    line = {
        "impl": "reference", "metric": metric_name(pname, args.mode), "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sample / value, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": workload_text(pname, B, args.mode, args.scaling, args.total, args.gpus) +
                                        f"; bounded sample of {sample} per step on the host CPU", "params": pname, "batch_per_step": sample},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port",
                         "sample": f"{sample} items per step, {args.steps} timed steps, std::thread static partition over {cores} threads",
                         "single_thread_ms_per_bootstrap": single_ms},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "outputs_correct": ok, "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--params", default=PARAMS, choices=["80", "110", "128", "uint4"], help="parameter set (uint4: LUT bootstrap workload)")
    ap.add_argument("--mode", default=None, choices=["fast", "exact"], help="transform mode (default: fast; exact for uint4)")
    ap.add_argument("--batch", type=int, default=0, help="units per GPU per step (default: 65,536 gates / 32,768 LUT bootstraps)")
    ap.add_argument("--total", type=int, default=0, help="units per step over all GPUs (with --scaling strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--kct", type=int, default=0)
    ap.add_argument("--no-tma", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pageable", action="store_true", help="e2e leg from pageable host memory (what a Zig caller's page_allocator gives)")
    ap.add_argument("--host-copy-threads", type=int, default=-1, help="tuning key host_copy_threads for the e2e leg (-1: library default)")
    ap.add_argument("--host-pipeline", type=int, default=-1, help="tuning key host_pipeline for the e2e leg (-1: library default)")
    args = ap.parse_args()
    if args.mode is None:
        args.mode = "exact" if args.params.startswith("uint") else "fast"
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    import tfhe_b200
    from tfhe_b200 import hostkeys as HK

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pname = args.params
    params = tfhe_b200.PARAM_SETS[pname]
    is_lut = pname.startswith("uint")
    from tfhe_b200 import dist as D
    if args.scaling == "strong":
        total = args.total or (1 << 20)
        lo, hi = D.shard_range(total, rank, world)
        B = hi - lo
    else:
        B = args.batch or (LUT_BATCH if is_lut else BATCH)
        total = B * world
    w = params.n + 1
    FLOP_PER_BOOTSTRAP, BSK_BYTES, KSK_BYTES = work_per_bootstrap(params)

    # ---- keys: generated on rank 0, one NCCL broadcast, re-laid-out on device by the library
    ctx = tfhe_b200.Context(params, devices=[local_rank])
    sk = ck = None
    if rank == 0:
        sk, ck = HK.gen_cloud_key(params, seed=1)
    if world > 1:
        secret = np.concatenate([sk.key_lv0, sk.key_lv1]) if rank == 0 else None
        d_bsk, d_ksk, d_sec = D.broadcast_cloud_key(params, ck, secret, dev)
        torch.cuda.synchronize()
        ctx.load_key_device(0, d_bsk.data_ptr(), d_ksk.data_ptr(), HK.gen_decomposition_offset(params))
        if rank != 0:
            sec = d_sec.cpu().numpy().view(np.uint32)
            sk = HK.SecretKey(sec[: params.n].copy(), sec[params.n:].copy())
        del d_bsk, d_ksk, d_sec
        torch.cuda.empty_cache()
    else:
        ctx.load_cloud_key(ck)
    if args.mode == "exact":
        ctx.set_mode(tfhe_b200.MODE_EXACT)

    # ---- synthetic inputs (this rank's shard), pinned on the host and resident on the device
    work = make_work(pname, params, sk, B, seed=42 + rank)
    pin = (lambda t: t) if args.pageable else (lambda t: t.pin_memory())
    if is_lut:
        h_a = pin(torch.from_numpy(work["ct"].view(np.int32)))
        h_tv = pin(torch.from_numpy(work["tv"].view(np.int32)))
        d_a, d_tv = h_a.to(dev), h_tv.to(dev)
        h_b = h_ops = d_b = d_ops = None
    else:
        h_a = pin(torch.from_numpy(work["ca"].view(np.int32)))
        h_b = pin(torch.from_numpy(work["cb"].view(np.int32)))
        h_ops = pin(torch.from_numpy(work["ops"]))
        d_a, d_b, d_ops = h_a.to(dev), h_b.to(dev), h_ops.to(dev)
    h_out = pin(torch.empty((B, w), dtype=torch.int32))
    d_out = torch.empty((B, w), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    if args.kct:
        ctx.set_tuning("kct", args.kct)
    if args.no_tma:
        ctx.set_tuning("use_tma", 0)
    ctx.set_tuning("timing", 1)
    stream = torch.cuda.ExternalStream(ctx.stream(0), device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    def step_device(n=B):
        if is_lut:
            ctx.bootstrap_batch_device(0, d_a.data_ptr(), d_out.data_ptr(), n, d_tv.data_ptr(), False)
        else:
            ctx.gate_batch_device(0, 0, d_ops.data_ptr(), d_a.data_ptr(), d_b.data_ptr(), d_out.data_ptr(), n)

    fp64_peak = ctx.measure_fp64_tflops(0)

    # ---- device-resident throughput (value)
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launch_count()
    k1_ms, k2_ms = [], []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    ctx.sync()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    # per-kernel times of one more (untimed-for-value) step, from CUDA events on the launching stream
    for _ in range(3):
        step_device()
        ctx.sync()
        k1_ms.append(ctx.last_kernel_ms(0, 0)); k2_ms.append(ctx.last_kernel_ms(0, 1))
    out_dev = d_out.cpu().numpy().view(np.uint32)
    ok_dev = True if is_lut else bool((HK.decrypt_bools(out_dev, sk) == work["truth"]).all())

    # ---- end to end through the host-buffer C ABI (H2D + kernels + D2H per step); per-kernel event timing off: this is
    # the call exactly as a user makes it
    ctx.set_tuning("timing", 0)
    if args.host_copy_threads >= 0:
        ctx.set_tuning("host_copy_threads", args.host_copy_threads)
    if args.host_pipeline >= 0:
        ctx.set_tuning("host_pipeline", args.host_pipeline)
    out_np = h_out.numpy().view(np.uint32)
    if is_lut:
        a_np, tv_np = h_a.numpy().view(np.uint32), h_tv.numpy().view(np.uint32)
        step_host = lambda: ctx.bootstrap_batch(a_np, tv_np, out=out_np)
        h2d, d2h = B * w * 4 + 2 * 1024 * 4, B * w * 4
    else:
        a_np, b_np, ops_np = h_a.numpy().view(np.uint32), h_b.numpy().view(np.uint32), h_ops.numpy()
        step_host = lambda: ctx.gate_batch(ops_np, a_np, b_np, out=out_np)
        h2d, d2h = 2 * B * w * 4 + B * 4, B * w * 4
    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    ok_e2e = bool((out_np == out_dev).all()) and (is_lut or bool((HK.decrypt_bools(out_np, sk) == work["truth"]).all()))

    # ---- p50 latency of a single bootstrap (B = 1, device resident)
    lat = []
    for _ in range(15):
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        step_device(1)
        s1.record(stream)
        ctx.sync()
        lat.append(s0.elapsed_time(s1))
    lat_p50 = float(np.median(lat[3:]))

    # ---- max over ranks
    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max = float(t[0]), float(t[1])
    flags = torch.tensor([int(ok_dev), int(ok_e2e)], device=dev)
    if world > 1:
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)

    if rank == 0:
        total_units = total * args.steps
        value = total_units / (dev_ms_max * 1e-3)
        e2e = total_units / (e2e_ms_max * 1e-3)
        k1 = float(np.mean(k1_ms)); k2 = float(np.mean(k2_ms))
        achieved_tflops = FLOP_PER_BOOTSTRAP * B / (k1 * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        sm_mhz = clocks.get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
        arith_peak = 148 * 64 * 2 * sm_mhz * 1e6 / 1e12      # 64 FP64 FMA lanes per SM per clock
        kct = args.kct or 6      # six ciphertexts per CTA is the full-wave configuration (blind_rotate.cu)
        waves = -(-B // (148 * kct))
        io_bytes = B * ((1 if is_lut else 2) * w * 4 + 4 + 4100)      # operands + opcode in, one lv1 sample out
        traffic, traffic_src = None, None                  # dram bytes per launch of K1 from the committed ncu capture
        for name in ("r02_k1_traffic.json", "r01_k1_traffic.json"):
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", name)))
                if B == BATCH and pname == PARAMS and args.mode == "fast" and not args.kct:
                    traffic, traffic_src = tr["dram_bytes_total"], name
                break
            except Exception:
                continue
        roofline = {
            "kernel": "blind_rotate_exact_kernel" if args.mode == "exact" else "blind_rotate_kernel", "bound": "fp64",
            "achieved": achieved_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": achieved_tflops / fp64_peak if fp64_peak > 0 else None, "traffic": traffic,
            "traffic_source": f"profiles/{traffic_src} (ncu dram__bytes_read.sum + dram__bytes_write.sum, one launch of this workload; a committed "
                              "capture, not measured in this run)" if traffic_src else None,
            "peak_source": "measured live by tfhe_b200_measure_fp64_tflops (DFMA microbenchmark); MEASURED_PEAKS.json has no FP64 figure",
            "peak_arithmetic": arith_peak, "frac_of_arithmetic_peak": achieved_tflops / arith_peak,
            "peak_arithmetic_source": f"148 SMs x 64 FP64 FMA lanes x 2 flop x {sm_mhz:.0f} MHz (median SM clock under load in this run)",
            "algorithmic_flop_per_bootstrap": FLOP_PER_BOOTSTRAP, "kernel_ms": k1, "kernel_share_of_step": k1 / (k1 + k2),
            "hbm": {"algorithmic_bytes": BSK_BYTES * waves + io_bytes, "achieved_gbs": (BSK_BYTES * waves + io_bytes) / (k1 * 1e-3) / 1e9,
                    "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                    "note": f"north_star's amortised-key term: each of the {waves} lock-step CTA waves ({148 * kct} bootstraps) streams the "
                            f"{BSK_BYTES / 1e6:.1f} MB key from HBM once and shares it through L2; far from the HBM bound (the kernel is FP64/shared-memory bound)"},
            "keyswitch": {"kernel_ms": k2, "algorithmic_bytes": KSK_BYTES * ((1 << params.basebit) - 1) // (1 << params.basebit) + B * (4100 + w * 4),
                          "achieved_gbs": (KSK_BYTES * ((1 << params.basebit) - 1) // (1 << params.basebit) + B * (4100 + w * 4)) / (k2 * 1e-3) / 1e9},
        }
        line = {
            "metric": metric_name(pname, args.mode), "value": value, "unit": "bootstraps/s" if is_lut else UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(pname, B, args.mode, args.scaling, total, world),
                       "params": pname, "mode": args.mode, "batch_per_gpu": B, "total_per_step": total,
                       "parallelism": f"batch-sharded x{world}, keys replicated (one NCCL broadcast at load)",
                       "host_memory": "pageable" if args.pageable else "pinned",
                       "l2": f"inputs ({h2d / 1e6:.0f} MB per step) larger than L2; keys ({(BSK_BYTES + KSK_BYTES * 3 // 4) / 1e6:.0f} MB) larger than L2"},
            "e2e": {"value": e2e, "unit": "bootstraps/s" if is_lut else UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "latency_ms_p50_single_gate": lat_p50,
            "roofline": roofline, "clocks": clocks, "outputs_correct": bool(flags[0].item() and flags[1].item()),
        }
        if world == 1 and not args.no_cpu_baseline:
            rate, cores, sample, out_cpu, single_ms = cpu_reference_rate(pname, (sk, ck), work)
            line["cpu_baseline"] = {"value": rate, "unit": line["unit"], "cores": cores, "kind": "port",
                                    "sample": f"first {sample} items of the same batch, oracle C++ port, {cores} std::threads",
                                    "single_thread_ms_per_bootstrap": single_ms,
                                    "published_single_thread_ms_per_gate": 36.73,
                                    "matches_gpu_bit_exact": bool((out_cpu == out_dev[:sample]).all())}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
