#!/bin/bash
mkdir -p gpurun_out
{
echo "== kct=6 generic (no L=3 instantiation)"; TFHE_B200_LIB=$PWD/zig-tfhe_b200/build/libtfhe_b200_nolt3.so python tools/prof_one.py 6 $((148*6*6)) 3 latency_mode=0 | tail -2
echo "== kct=6 L=3 instantiation"; python tools/prof_one.py 6 $((148*6*6)) 3 latency_mode=0 | tail -2
} > gpurun_out/y_lt.log 2>&1; cat gpurun_out/y_lt.log
