#!/bin/bash
mkdir -p gpurun_out
for cfg in "pageable_t0:--pageable --host-copy-threads 0" "pageable_t1:--pageable --host-copy-threads 1" "pageable_t4:--pageable --host-copy-threads 4" "pageable_t8:--pageable --host-copy-threads 8"; do
  name=${cfg%%:*}; flags=${cfg#*:}
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline $flags > gpurun_out/n_bench_$name.json 2> gpurun_out/n_bench_$name.err
done
python - <<'PY' | tee gpurun_out/n_host_path.log
import json
for f in ("pageable_t0","pageable_t1","pageable_t4","pageable_t8"):
    try:
        d=json.load(open(f"gpurun_out/n_bench_{f}.json"))
        print(f, "device", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["outputs_correct"])
    except Exception as e: print(f,"failed",e)
PY
