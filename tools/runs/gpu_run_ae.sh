#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 3 --warmup 3 --params uint4 --mode fast --no-cpu-baseline > gpurun_out/ae_bench_uint4_fast.json 2> gpurun_out/ae_bench_uint4_fast.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/ae_bench_uint4_fast.json")); print("uint4 fast", round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"],3), d["outputs_correct"])
PY
( timeout 600 python -m pytest tests/test_gpu_exact_and_uint.py -m gpu -x -q ) 2>&1 | tail -2
