#!/bin/bash
mkdir -p gpurun_out
{
echo "== baseline kct=4"; python tools/prof_one.py 4 $((148*4*8)) 3 latency_mode=0
echo "== twt kct=4"; python tools/prof_one.py 4 $((148*4*8)) 3 latency_mode=0 twt=1
echo "== twt kct=6"; python tools/prof_one.py 6 $((148*6*6)) 3 latency_mode=0 twt=1
echo "== twt kct=5"; python tools/prof_one.py 5 $((148*5*6)) 3 latency_mode=0 twt=1
echo "== old kct=6"; python tools/prof_one.py 6 $((148*6*6)) 3 latency_mode=0
} > gpurun_out/e_twt.log 2>&1
grep -E "==|K1 ms|ok|rror" gpurun_out/e_twt.log
