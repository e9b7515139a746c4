#!/bin/bash
# config 4 only (UINT4 LUT, exact mode) on N GPUs: tools/scale_uint4.sh N
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" -gt 1 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --params uint4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_uint4_exact_${N}gpu.json 2> gpurun_out/r02b_uint4_exact_${N}gpu.err
else
  timeout 600 python bench.py --gpus 1 --params uint4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_uint4_exact_${N}gpu.json 2> gpurun_out/r02b_uint4_exact_${N}gpu.err
fi
python - "$N" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/r02b_uint4_exact_{sys.argv[1]}gpu.json").read().strip().splitlines()[-1])
print("uint4_exact N=" + sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ok", d["outputs_correct"], "ms/step", round(d["ms_per_step"], 1))
PY
