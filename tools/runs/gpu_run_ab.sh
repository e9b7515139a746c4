#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_full_configs.py tests/test_gpu_parity.py -m gpu -x -q ) 2>&1 | tail -3
timeout 300 python bench.py --steps 4 --warmup 3 --pageable --no-cpu-baseline > gpurun_out/ab_bench_pageable.json 2> gpurun_out/ab_bench_pageable.err
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ab_bench_pinned.json 2> gpurun_out/ab_bench_pinned.err
python - <<'PY' | tee gpurun_out/ab_pipeline.log
import json
for f in ("ab_bench_pageable","ab_bench_pinned"):
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, "device", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["outputs_correct"])
PY
