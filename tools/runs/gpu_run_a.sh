#!/bin/bash
# round-2 GPU call A: full GPU test suite at HEAD, bench (default + the new flags), compute-sanitizer on the tiny set
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/a_gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/a_pytest.log 2>&1
tail -5 gpurun_out/a_pytest.log
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/a_bench_default.json 2> gpurun_out/a_bench_default.err
tail -c 600 gpurun_out/a_bench_default.json
timeout 300 python bench.py --steps 2 --warmup 3 --params uint4 > gpurun_out/a_bench_uint4.json 2> gpurun_out/a_bench_uint4.err
tail -c 400 gpurun_out/a_bench_uint4.json
timeout 300 python bench.py --steps 2 --warmup 3 --params 80 --no-cpu-baseline > gpurun_out/a_bench_80.json 2> gpurun_out/a_bench_80.err
for tool in memcheck racecheck synccheck; do
  ( time timeout 420 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_driver.py ) > gpurun_out/a_sanitize_$tool.log 2>&1
  tail -4 gpurun_out/a_sanitize_$tool.log
done
