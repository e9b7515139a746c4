#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/o_pytest.log 2>&1
tail -15 gpurun_out/o_pytest.log
