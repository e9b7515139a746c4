#!/bin/bash
mkdir -p gpurun_out
{ echo "== twiddle fetch issued before the exchange: K1 kct=6"; python tools/prof_one.py 6 $((148*6*6)) 4 latency_mode=0 | tail -3; } > gpurun_out/aj_prefetch.log 2>&1; cat gpurun_out/aj_prefetch.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
