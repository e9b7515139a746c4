#!/bin/bash
mkdir -p gpurun_out
{
echo "== diag build, last-arriver: 0 32 1"; for m in 0 32 1; do TFHE_B200_LIB=$PWD/zig-tfhe_b200/build/libtfhe_b200_diag.so python tools/prof_one.py 4 $((148*4*6)) 2 latency_mode=0 diag=$m | grep "K1 ms" | tail -1; done
echo "== diag build, kct=6 twt: 0 32 1"; for m in 0 32 1; do TFHE_B200_LIB=$PWD/zig-tfhe_b200/build/libtfhe_b200_diag.so python tools/prof_one.py 6 $((148*6*4)) 2 latency_mode=0 twt=1 diag=$m | grep "K1 ms" | tail -1; done
} > gpurun_out/h_ring.log 2>&1
cat gpurun_out/h_ring.log
