#!/bin/bash
# BASELINE configs on N GPUs of one box from ONE build: default weak-scaling line, config 5 (2^20 gates, strong scaling,
# 80/110/128-bit) and config 4 (UINT4 LUT, exact mode).  Usage: tools/scale_run.sh N [tag]
N=${1:-1}; TAG=${2:-r02}
mkdir -p gpurun_out
run() {  # name, bench flags...
  local name=$1; shift
  if [ "$N" -gt 1 ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" \
      > gpurun_out/${TAG}_${name}_${N}gpu.json 2> gpurun_out/${TAG}_${name}_${N}gpu.err
  else
    timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/${TAG}_${name}_${N}gpu.json 2> gpurun_out/${TAG}_${name}_${N}gpu.err
  fi
  python - "$name" "$N" "gpurun_out/${TAG}_${name}_${N}gpu.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[3]).read().strip().splitlines()[-1])
    print(sys.argv[1], "N=" + sys.argv[2], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ok", d["outputs_correct"], "ms/step", round(d["ms_per_step"], 1))
except Exception as e:
    print(sys.argv[1], "N=" + sys.argv[2], "FAILED", e)
PY
}
run weak128 --steps 5 --warmup 3 --no-cpu-baseline
for p in 80 110 128; do run strong1M_$p --params $p --total 1048576 --scaling strong --steps 2 --warmup 3 --no-cpu-baseline; done
run uint4_exact --params uint4 --steps 3 --warmup 3 --no-cpu-baseline
