#!/bin/bash
mkdir -p gpurun_out
python tools/prof_one.py 4 1184 2 latency_mode=0 > gpurun_out/d_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:blind_rotate_kernel -s 1 -c 1 -o gpurun_out/d_k1 python tools/prof_one.py 4 1184 2 latency_mode=0 > gpurun_out/d_ncu.log 2>&1
tail -3 gpurun_out/d_plain.log; tail -3 gpurun_out/d_ncu.log
