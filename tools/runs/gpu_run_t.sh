#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/adder_bench.py > gpurun_out/t_adder.log 2>&1; tail -8 gpurun_out/t_adder.log
timeout 300 python tools/stress.py 40 3 > gpurun_out/t_stress.log 2>&1; tail -3 gpurun_out/t_stress.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/t_bench_reference.json 2>gpurun_out/t_bench_reference.err; cut -c1-300 gpurun_out/t_bench_reference.json
