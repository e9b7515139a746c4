#!/bin/bash
mkdir -p gpurun_out
{
echo "== kct=6 twt + X1 alias (double-buffered X2)"; python tools/prof_one.py 6 $((148*6*6)) 3 latency_mode=0 | tail -3
echo "== kct=5 twt"; python tools/prof_one.py 5 $((148*5*6)) 3 latency_mode=0 twt=1 | tail -2
} > gpurun_out/l_alias.log 2>&1
cat gpurun_out/l_alias.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
