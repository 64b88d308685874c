#!/bin/bash
# GPU call Q (1 GPU): final round-2 evidence — launch list of the bench command, ncu --set full of the fp16x3 hop.
mkdir -p gpurun_out
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/q_bench_under_ncu.log 2>&1; python tools/launch_summary.py gpurun_out/r2_launches.csv > gpurun_out/r2_launches_summary.txt 2>&1; head -5 gpurun_out/r2_launches_summary.txt
echo "== ncu full hop16"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_rbu_tc16 -s 2 -c 1 -o gpurun_out/r2_prof_hop16 python tools/profile_tc16.py 16 > gpurun_out/q_ncu_hop16.log 2>&1; tail -1 gpurun_out/q_ncu_hop16.log
