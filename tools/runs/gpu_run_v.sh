#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/v_pytest.log 2>&1; tail -4 gpurun_out/v_pytest.log
timeout 600 python tools/stress.py 120 7 > gpurun_out/v_stress.log 2>&1; tail -2 gpurun_out/v_stress.log
timeout 300 python bench.py > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; tail -c 700 gpurun_out/v_bench.json
timeout 200 python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -1
