#!/bin/bash
# end of round 2: ncu --set full of the final K1 (paired inverse transforms), then the full suite, bench and smoke on the same build
mkdir -p gpurun_out
python tools/prof_one.py 6 1776 2 latency_mode=0 > gpurun_out/ak_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:blind_rotate_kernel -s 1 -c 1 -o gpurun_out/ak_k1 python tools/prof_one.py 6 1776 2 latency_mode=0 > gpurun_out/ak_ncu.log 2>&1
tail -3 gpurun_out/ak_plain.log; tail -3 gpurun_out/ak_ncu.log
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/ak_pytest.log 2>&1; tail -4 gpurun_out/ak_pytest.log
timeout 300 python bench.py > gpurun_out/ak_bench.json 2> gpurun_out/ak_bench.err; tail -c 400 gpurun_out/ak_bench.json
timeout 200 python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -1
