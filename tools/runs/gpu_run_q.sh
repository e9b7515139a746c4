#!/bin/bash
mkdir -p gpurun_out
{
echo "== kct=6 twt=2 (key through TMEM), small"; timeout 120 python tools/prof_one.py 6 24 1 latency_mode=0 twt=2 | tail -2
echo "== kct=6 twt=2 (key through TMEM)"; timeout 120 python tools/prof_one.py 6 $((148*6*6)) 3 latency_mode=0 twt=2 | tail -3
echo "== kct=6 twt=1"; timeout 120 python tools/prof_one.py 6 $((148*6*6)) 3 latency_mode=0 twt=1 | tail -2
} > gpurun_out/q_ktm.log 2>&1
cat gpurun_out/q_ktm.log
