#!/bin/bash
mkdir -p gpurun_out
{
echo "== kct=6 L=3 instantiation, compile-time offset"; python tools/prof_one.py 6 $((148*6*6)) 4 latency_mode=0 | tail -3
echo "== kct=4"; python tools/prof_one.py 4 $((148*4*8)) 3 latency_mode=0 | tail -2
} > gpurun_out/aa_offset.log 2>&1; cat gpurun_out/aa_offset.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
