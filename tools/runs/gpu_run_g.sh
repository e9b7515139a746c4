#!/bin/bash
mkdir -p gpurun_out
{
echo "== last-arriver ring kct=4"; python tools/prof_one.py 4 $((148*4*8)) 3 latency_mode=0 | tail -3
echo "== last-arriver ring kct=2"; python tools/prof_one.py 2 $((148*2*8)) 3 latency_mode=0 | tail -3
echo "== last-arriver ring kct=6 twt"; python tools/prof_one.py 6 $((148*6*6)) 3 latency_mode=0 twt=1 | tail -3
echo "== last-arriver ring kct=4 twt"; python tools/prof_one.py 4 $((148*4*8)) 3 latency_mode=0 twt=1 | tail -3
echo "== poll ring kct=4"; TFHE_B200_LIB=$PWD/zig-tfhe_b200/build/libtfhe_b200_poll.so python tools/prof_one.py 4 $((148*4*8)) 3 latency_mode=0 | tail -3
echo "== diag build, last-arriver: 0 1 2 3"; for m in 0 1 2 3; do TFHE_B200_LIB=$PWD/zig-tfhe_b200/build/libtfhe_b200_diag.so python tools/prof_one.py 4 $((148*4*6)) 2 latency_mode=0 diag=$m | grep "K1 ms" | tail -1; done
} > gpurun_out/g_ring.log 2>&1
cat gpurun_out/g_ring.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
