#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_exact_and_uint.py tests/test_gpu_full_configs.py -m gpu -x -q -k "tensor_core or keyswitch or uint or config4" ) 2>&1 | tail -6
timeout 300 python bench.py --steps 3 --warmup 3 --params uint4 > gpurun_out/w_bench_uint4.json 2> gpurun_out/w_bench_uint4.err
python - <<'PY' | tee gpurun_out/w_k2t_general.log
import json
d=json.load(open("gpurun_out/w_bench_uint4.json"))
print("uint4 exact: value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "K1x ms", round(d["roofline"]["kernel_ms"],1), "K2 ms", round(d["roofline"]["keyswitch"]["kernel_ms"],2), "ok", d["outputs_correct"], d["cpu_baseline"]["matches_gpu_bit_exact"])
PY
