#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_exact_and_uint.py tests/test_gpu_full_configs.py -m gpu -x -q -k "exact or uint4 or config4" ) 2>&1 | tail -3
timeout 300 python bench.py --steps 3 --warmup 3 --params uint4 > gpurun_out/z_bench_uint4.json 2> gpurun_out/z_bench_uint4.err
timeout 300 python bench.py --steps 2 --warmup 3 --params 128 --mode exact --batch 16384 --no-cpu-baseline > gpurun_out/z_bench_128_exact.json 2> gpurun_out/z_bench_128_exact.err
python - <<'PY' | tee gpurun_out/z_k1x_lt.log
import json
for f in ("z_bench_uint4","z_bench_128_exact"):
    d=json.load(open(f"gpurun_out/{f}.json"))
    print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "K1x ms", round(d["roofline"]["kernel_ms"],1), "K2 ms", round(d["roofline"]["keyswitch"]["kernel_ms"],2), "ok", d["outputs_correct"], d.get("cpu_baseline",{}).get("matches_gpu_bit_exact"))
PY
