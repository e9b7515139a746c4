#!/bin/bash
mkdir -p gpurun_out
export TFHE_B200_LIB=$PWD/zig-tfhe_b200/build/libtfhe_b200_diag.so
{
for m in 0 1 2 3 4 8 16 24 28 31; do
echo "== diag=$m kct=4"; python tools/prof_one.py 4 $((148*4*6)) 2 latency_mode=0 diag=$m | grep "K1 ms" | tail -1
done
for m in 0 1 2 24 31; do
echo "== diag=$m kct=6 twt"; python tools/prof_one.py 6 $((148*6*4)) 2 latency_mode=0 twt=1 diag=$m | grep "K1 ms" | tail -1
done
} > gpurun_out/f_diag.log 2>&1
cat gpurun_out/f_diag.log
