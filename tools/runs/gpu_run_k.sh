#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/k_pytest.log 2>&1
tail -5 gpurun_out/k_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/k_bench.json"))
print(round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"],3), round(d["roofline"]["kernel_ms"],1), round(d["roofline"]["keyswitch"]["kernel_ms"],2), d["outputs_correct"], d["latency_ms_p50_single_gate"], d["cpu_baseline"]["matches_gpu_bit_exact"], d["gpu_launches"])
PY
python tools/latency.py > gpurun_out/k_latency.log 2>&1; tail -12 gpurun_out/k_latency.log
