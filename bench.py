#!/usr/bin/env python
"""bench.py -- bootstrapped gates/s of the zig-tfhe gate-bootstrapping hot path on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W]             (N = 1)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...                            (host-CPU arm)

Workload (BASELINE.json configs[1]): 65,536 independent AND/XOR gates (half each) at
SECURITY_128_BIT per GPU ("weak" scaling: every rank owns its own 65,536-gate shard; keys are
generated once on rank 0 and broadcast with one NCCL broadcast at key load; there is no collective
on the hot path).  One step = one pass of the whole path (gate linear part -> blind rotation ->
sample extract -> key switch) over the batch.

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
METRIC = "bootstrapped gates/sec (whole box, 128-bit)"
UNIT = "gates/s"
# algorithmic work per bootstrap, SURVEY.md section 8d / BASELINE.md section 3 (128-bit set)
FLOP_PER_BOOTSTRAP = 180_633_600
BSK_BYTES = 68_812_800
KSK_BYTES = 103_366_656
AND, XOR = 2, 3


def synth_inputs(params, sk, batch, seed):
    from tfhe_b200 import hostkeys as HK
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 2, batch).astype(np.uint8)
    b = rng.integers(0, 2, batch).astype(np.uint8)
    ca = HK.encrypt_bools(a, params, sk, rng)
    cb = HK.encrypt_bools(b, params, sk, rng)
    ops = np.where(np.arange(batch) < batch // 2, AND, XOR).astype(np.int32)
    truth = np.where(ops == AND, a & b, a ^ b).astype(np.uint8)
    return ca, cb, ops, truth


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


def cpu_reference_rate(keys_tuple, ca, cb, ops, seconds_target=12.0):
    """times the CPU oracle (port of the reference path) with all host threads on a bounded sample"""
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    sk, ck = keys_tuple
    orc = O.Oracle(PARAMS)
    keys = O.Keys(sk.key_lv0, sk.key_lv1, ck.bootstrapping_key, ck.key_switching_key, ck.decomposition_offset, ck.blind_rotate_testvec)
    cores = O.hardware_threads()
    probe = min(len(ca), 2 * cores)
    t0 = time.perf_counter()
    orc.gate_batch(ops[:probe], ca[:probe], cb[:probe], keys, nthreads=cores)
    per_round = (time.perf_counter() - t0) / 2.0          # seconds for `cores` gates
    sample = int(min(len(ca), max(4 * cores, cores * max(1, int(seconds_target / max(per_round, 1e-3))))))
    t0 = time.perf_counter()
    out = orc.gate_batch(ops[:sample], ca[:sample], cb[:sample], keys, nthreads=cores)
    dt = time.perf_counter() - t0
    return sample / dt, cores, sample, out


def run_reference(args):
    """--impl reference: the reference's own algorithm on the host cores (oracle port; Zig cannot be built here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import tfhe_b200
    from tfhe_b200 import hostkeys as HK
    params = tfhe_b200.PARAM_SETS[PARAMS]
    sk, ck = HK.gen_cloud_key(params, seed=1)
    sample_max = 4096
    ca, cb, ops, truth = synth_inputs(params, sk, sample_max, seed=42)
    rates = []
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    orc = O.Oracle(PARAMS)
    keys = O.Keys(sk.key_lv0, sk.key_lv1, ck.bootstrapping_key, ck.key_switching_key, ck.decomposition_offset, ck.blind_rotate_testvec)
    cores = O.hardware_threads()
    sample = int(min(sample_max, max(2 * cores, 16 * cores)))
    total = args.warmup + args.steps
    budget = 150.0
    t_probe = time.perf_counter()
    orc.gate_batch(ops[:cores], ca[:cores], cb[:cores], keys, nthreads=cores)
    per_round = time.perf_counter() - t_probe
    rounds = max(1, min(16, int(budget / total / max(per_round, 1e-3))))
    sample = min(sample_max, rounds * cores)
    ok = True
    for step in range(total):
        t0 = time.perf_counter()
        out = orc.gate_batch(ops[:sample], ca[:sample], cb[:sample], keys, nthreads=cores)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            rates.append(sample / dt)
        ok = ok and bool((HK.decrypt_bools(out, sk) == truth[:sample]).all())
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sample / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": f"AND/XOR gates at SECURITY_128_BIT, bounded sample of {sample} gates per step (of {BATCH})",
                                        "params": PARAMS, "batch_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} gates per step, {args.steps} timed steps, std::thread static partition over {cores} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "outputs_correct": ok, "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="gates per GPU per step (default: the BASELINE configuration)")
    ap.add_argument("--kct", type=int, default=0)
    ap.add_argument("--no-tma", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
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
    params = tfhe_b200.PARAM_SETS[PARAMS]
    B = args.batch
    w = params.n + 1

    # ---- keys: generated on rank 0, one NCCL broadcast, re-laid-out on device by the library
    from tfhe_b200 import dist as D
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

    # ---- synthetic inputs (this rank's shard), pinned on the host and resident on the device
    ca, cb, ops, truth = synth_inputs(params, sk, B, seed=42 + rank)
    h_a = torch.from_numpy(ca.view(np.int32)).pin_memory()
    h_b = torch.from_numpy(cb.view(np.int32)).pin_memory()
    h_ops = torch.from_numpy(ops).pin_memory()
    h_out = torch.empty((B, w), dtype=torch.int32).pin_memory()
    d_a, d_b, d_ops = h_a.to(dev), h_b.to(dev), h_ops.to(dev)
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

    def step_device():
        ctx.gate_batch_device(0, 0, d_ops.data_ptr(), d_a.data_ptr(), d_b.data_ptr(), d_out.data_ptr(), B)

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
    ok_dev = bool((HK.decrypt_bools(out_dev, sk) == truth).all())

    # ---- end to end through the host-buffer C ABI (H2D + kernels + D2H per step); per-kernel event timing off: this is
    # the call exactly as a user makes it
    ctx.set_tuning("timing", 0)
    a_np, b_np, ops_np, out_np = h_a.numpy().view(np.uint32), h_b.numpy().view(np.uint32), h_ops.numpy(), h_out.numpy().view(np.uint32)
    ctx.gate_batch(ops_np, a_np, b_np, out=out_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.gate_batch(ops_np, a_np, b_np, out=out_np)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    ok_e2e = bool((HK.decrypt_bools(out_np, sk) == truth).all()) and bool((out_np == out_dev).all())

    # ---- p50 latency of a single bootstrapped gate (B = 1, device resident)
    lat = []
    for _ in range(15):
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        ctx.gate_batch_device(0, 0, d_ops.data_ptr(), d_a.data_ptr(), d_b.data_ptr(), d_out.data_ptr(), 1)
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
        total_gates = B * world * args.steps
        value = total_gates / (dev_ms_max * 1e-3)
        e2e = total_gates / (e2e_ms_max * 1e-3)
        k1 = float(np.mean(k1_ms)); k2 = float(np.mean(k2_ms))
        achieved_tflops = FLOP_PER_BOOTSTRAP * B / (k1 * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        kct = args.kct or 4
        waves = -(-B // (148 * kct))
        io_bytes = B * (2 * w * 4 + 4 + 4100)           # two operands + opcode in, one lv1 sample out
        traffic = None                                  # dram bytes per launch of K1 from the committed ncu capture
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_k1_traffic.json")))
            if B == BATCH and not args.kct:
                traffic = tr["dram_bytes_total"]
        except Exception:
            pass
        roofline = {
            "kernel": "blind_rotate_kernel", "bound": "fp64", "achieved": achieved_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": achieved_tflops / fp64_peak if fp64_peak > 0 else None, "traffic": traffic,
            "traffic_source": "profiles/r01_k1_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, one launch of this workload)",
            "peak_source": "measured live by tfhe_b200_measure_fp64_tflops (DFMA microbenchmark); MEASURED_PEAKS.json has no FP64 figure",
            "algorithmic_flop_per_bootstrap": FLOP_PER_BOOTSTRAP, "kernel_ms": k1, "kernel_share_of_step": k1 / (k1 + k2),
            "hbm": {"algorithmic_bytes": BSK_BYTES * waves + io_bytes, "achieved_gbs": (BSK_BYTES * waves + io_bytes) / (k1 * 1e-3) / 1e9,
                    "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                    "note": f"north_star's amortised-key term: each of the {waves} lock-step CTA waves ({148 * kct} bootstraps) streams the "
                            "68.8 MB key from HBM once and shares it through L2; far from the HBM bound (the kernel is FP64/shared-memory bound)"},
            "keyswitch": {"kernel_ms": k2, "algorithmic_bytes": KSK_BYTES * 3 // 4 + B * (4100 + w * 4),
                          "achieved_gbs": (KSK_BYTES * 3 // 4 + B * (4100 + w * 4)) / (k2 * 1e-3) / 1e9},
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"{B} independent AND/XOR gates (half/half) at SECURITY_128_BIT per GPU, seeded keys and ciphertexts",
                       "params": PARAMS, "batch_per_gpu": B, "parallelism": f"batch-sharded x{world}, keys replicated (one NCCL broadcast at load)",
                       "l2": "inputs (367 MB per step) larger than L2; keys (172 MB) larger than L2"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(2 * B * w * 4 + B * 4), "d2h_bytes_per_step": int(B * w * 4)},
            "gpu_launches": int(launches), "latency_ms_p50_single_gate": lat_p50,
            "roofline": roofline, "clocks": clocks, "outputs_correct": bool(flags[0].item() and flags[1].item()),
        }
        if world == 1 and not args.no_cpu_baseline:
            rate, cores, sample, out_cpu = cpu_reference_rate((sk, ck), ca, cb, ops)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"first {sample} gates of the same batch, oracle C++ port, {cores} std::threads",
                                    "matches_gpu_bit_exact": bool((out_cpu == out_dev[:sample]).all())}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
