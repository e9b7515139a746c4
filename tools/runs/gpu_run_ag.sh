#!/bin/bash
mkdir -p gpurun_out
L=$PWD/zig-tfhe_b200/build/libtfhe_b200_unrollh.so
{
echo "== h loop unrolled: K1 kct=6"; TFHE_B200_LIB=$L python tools/prof_one.py 6 $((148*6*6)) 3 latency_mode=0 | tail -2
echo "== h loop unrolled: exact uint4"; TFHE_B200_LIB=$L python tools/prof_exact.py $((148*6*6)) | tail -2
echo "== baseline: K1 kct=6"; python tools/prof_one.py 6 $((148*6*6)) 3 latency_mode=0 | tail -2
echo "== baseline: exact uint4"; python tools/prof_exact.py $((148*6*6)) | tail -2
} > gpurun_out/ag_unrollh.log 2>&1; cat gpurun_out/ag_unrollh.log
