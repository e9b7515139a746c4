#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_exact_and_uint.py tests/test_gpu_full_configs.py -m gpu -x -q ) > gpurun_out/b_pytest.log 2>&1
tail -15 gpurun_out/b_pytest.log
timeout 300 python bench.py --steps 3 --warmup 3 --params uint4 > gpurun_out/b_bench_uint4.json 2> gpurun_out/b_bench_uint4.err
tail -c 300 gpurun_out/b_bench_uint4.json; tail -3 gpurun_out/b_bench_uint4.err
timeout 300 python bench.py --steps 2 --warmup 3 --params uint4 --mode fast --no-cpu-baseline > gpurun_out/b_bench_uint4_fast.json 2> gpurun_out/b_bench_uint4_fast.err
timeout 300 python bench.py --steps 2 --warmup 3 --params 128 --mode exact --batch 16384 --no-cpu-baseline > gpurun_out/b_bench_128_exact.json 2> gpurun_out/b_bench_128_exact.err
python - <<'PY'
import json
for f in ("b_bench_uint4","b_bench_uint4_fast","b_bench_128_exact"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"],3), round(d["roofline"]["kernel_ms"],1), round(d["roofline"]["keyswitch"]["kernel_ms"],1), d["outputs_correct"], d.get("cpu_baseline",{}).get("matches_gpu_bit_exact"))
    except Exception as e: print(f, "failed", e)
PY
