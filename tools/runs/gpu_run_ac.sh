#!/bin/bash
mkdir -p gpurun_out
python tools/prof_exact.py 1776 > gpurun_out/ac_plain_ex.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:blind_rotate_exact_rb -s 1 -c 1 -o gpurun_out/ac_k1x python tools/prof_exact.py 1776 > gpurun_out/ac_ncu_ex.log 2>&1
tail -2 gpurun_out/ac_plain_ex.log
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/ac_pytest.log 2>&1; grep -E "passed|failed" gpurun_out/ac_pytest.log
timeout 300 python bench.py > gpurun_out/ac_bench.json 2> gpurun_out/ac_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/ac_bench.json")); print(round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"],3), d["outputs_correct"], d["cpu_baseline"]["matches_gpu_bit_exact"], d["latency_ms_p50_single_gate"])
PY
