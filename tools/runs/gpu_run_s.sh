#!/bin/bash
mkdir -p gpurun_out
{
echo "== kct=6 twt=3 (X1 through shuffles)"; timeout 120 python tools/prof_one.py 6 $((148*6*6)) 3 latency_mode=0 twt=3 | tail -3
echo "== kct=6 twt=1"; timeout 120 python tools/prof_one.py 6 $((148*6*6)) 3 latency_mode=0 twt=1 | tail -2
} > gpurun_out/s_x1s.log 2>&1
cat gpurun_out/s_x1s.log
