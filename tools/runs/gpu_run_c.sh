#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_core or keyswitch" ) > gpurun_out/c_pytest.log 2>&1
tail -25 gpurun_out/c_pytest.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
python - <<'PY'
import json
for f in ("c_bench",):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"],3), round(d["roofline"]["kernel_ms"],1), round(d["roofline"]["keyswitch"]["kernel_ms"],2), d["outputs_correct"])
    except Exception as e: print(f, "failed", e)
PY
tail -3 gpurun_out/c_bench.err
