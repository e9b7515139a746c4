#!/bin/bash
mkdir -p gpurun_out
{
echo "== kct=6 twt alias team=2"; python tools/prof_one.py 6 $((148*6*6)) 3 latency_mode=0 team=2 | tail -3
echo "== kct=6 twt alias team=1"; python tools/prof_one.py 6 $((148*6*6)) 3 latency_mode=0 | tail -2
} > gpurun_out/m_team.log 2>&1
cat gpurun_out/m_team.log
