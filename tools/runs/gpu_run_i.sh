#!/bin/bash
mkdir -p gpurun_out
D=$PWD/zig-tfhe_b200/build
{
echo "== diag: 0 64 32 1 (kct=4)"; for m in 0 64 32 1; do TFHE_B200_LIB=$D/libtfhe_b200_diag.so python tools/prof_one.py 4 $((148*4*6)) 2 latency_mode=0 diag=$m | grep "K1 ms" | tail -1; done
echo "== 4 stages kct=4"; TFHE_B200_LIB=$D/libtfhe_b200_st4.so python tools/prof_one.py 4 $((148*4*8)) 3 latency_mode=0 | tail -2
echo "== 4 stages kct=3"; TFHE_B200_LIB=$D/libtfhe_b200_st4.so python tools/prof_one.py 3 $((148*3*8)) 3 latency_mode=0 | tail -2
echo "== 3 stages kct=3"; python tools/prof_one.py 3 $((148*3*8)) 3 latency_mode=0 | tail -2
echo "== team2 kct=4"; python tools/prof_one.py 4 $((148*4*8)) 3 latency_mode=0 team=2 | tail -2
echo "== team2 kct=6"; python tools/prof_one.py 6 $((148*6*6)) 3 latency_mode=0 team=2 | tail -2
echo "== kct=6 twt"; python tools/prof_one.py 6 $((148*6*6)) 3 latency_mode=0 twt=1 | tail -2
} > gpurun_out/i_ring.log 2>&1
cat gpurun_out/i_ring.log
