#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 3 --warmup 3 --params uint4 --mode fast --no-cpu-baseline > gpurun_out/x_bench_uint4_fast.json 2> gpurun_out/x_bench_uint4_fast.err
timeout 300 python bench.py --steps 3 --warmup 3 --pageable --no-cpu-baseline > gpurun_out/x_bench_pageable.json 2> gpurun_out/x_bench_pageable.err
python tools/latency.py > gpurun_out/x_latency.log 2>&1
python - <<'PY'
import json
for f in ("x_bench_uint4_fast","x_bench_pageable"):
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"],3), d["outputs_correct"])
PY
grep "latency_mode=2" gpurun_out/x_latency.log | head -8
