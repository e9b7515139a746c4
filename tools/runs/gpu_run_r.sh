#!/bin/bash
# evidence: ncu --set full of K2t and the exact kernel, DRAM bytes of K1 at the bench workload, launch list of the bench command
mkdir -p gpurun_out
python tools/prof_one.py 0 4096 2 > gpurun_out/r_plain_k2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:keyswitch_tc_kernel -s 1 -c 1 -o gpurun_out/r_k2t python tools/prof_one.py 0 4096 2 > gpurun_out/r_ncu_k2.log 2>&1
python tools/prof_exact.py 1184 > gpurun_out/r_plain_ex.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:blind_rotate_exact_rb -s 1 -c 1 -o gpurun_out/r_k1x python tools/prof_exact.py 1184 > gpurun_out/r_ncu_ex.log 2>&1
python tools/prof_one.py 0 65536 2 > gpurun_out/r_plain_traffic.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex.sum --clock-control none -k regex:"blind_rotate|keyswitch_tc" --csv --log-file gpurun_out/r_traffic.csv python tools/prof_one.py 0 65536 2 > gpurun_out/r_ncu_traffic.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r_bench_plain.json 2> gpurun_out/r_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r_bench_ncu.log 2>&1
ls -la gpurun_out | grep " r_"
tail -2 gpurun_out/r_plain_k2.log gpurun_out/r_plain_ex.log
