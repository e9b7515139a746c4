#!/bin/bash
mkdir -p gpurun_out
{
echo "== split-phase LA kct=4"; python tools/prof_one.py 4 $((148*4*8)) 3 latency_mode=0 | tail -2
echo "== split-phase LA kct=6 twt"; python tools/prof_one.py 6 $((148*6*6)) 3 latency_mode=0 twt=1 | tail -2
echo "== split-phase LA kct=4 twt"; python tools/prof_one.py 4 $((148*4*8)) 3 latency_mode=0 twt=1 | tail -2
echo "== split-phase LA kct=3"; python tools/prof_one.py 3 $((148*3*8)) 3 latency_mode=0 | tail -2
} > gpurun_out/j_ring.log 2>&1
cat gpurun_out/j_ring.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
