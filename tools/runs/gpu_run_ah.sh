#!/bin/bash
mkdir -p gpurun_out
{ echo "== digit loop rolled (constants still compile-time): K1 kct=6"; TFHE_B200_LIB=$PWD/zig-tfhe_b200/build/libtfhe_b200_rolll.so python tools/prof_one.py 6 $((148*6*6)) 3 latency_mode=0 | tail -2; } > gpurun_out/ah_rolll.log 2>&1; cat gpurun_out/ah_rolll.log
