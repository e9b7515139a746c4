#!/bin/bash
# Builds a variant of the library with extra -D flags: tools/build_variant.sh NAME "-DTFHE_B200_DIAG ..."
#   -> zig-tfhe_b200/build/libtfhe_b200_NAME.so (git-ignored; select it with TFHE_B200_LIB=<path>)
# Flags: TFHE_B200_DIAG (switchable parts of the throughput kernel, see blind_rotate.cu), TFHE_B200_RING_POLL (round-1 key-ring refill)
set -e
NAME=$1; DEFS=$2
cd "$(dirname "$0")/../zig-tfhe_b200"
mkdir -p build/$NAME
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --fmad=false -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -diag-suppress 177 $DEFS"
for f in blind_rotate blind_rotate_exact keyswitch keyswitch_tc key_layout keygen key_file capi; do
  nvcc $FLAGS -c -o build/$NAME/$f.o csrc/$f.cu &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/libtfhe_b200_$NAME.so build/$NAME/*.o
echo built build/libtfhe_b200_$NAME.so
