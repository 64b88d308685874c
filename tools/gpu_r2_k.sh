#!/bin/bash
# GPU call K (1 GPU): where does the fp16x3 scan's step go?  W-ring depth sensitivity + ncu stall samples.
mkdir -p gpurun_out
for v in r16s3 r16s4; do echo "== stages ${v#r16s}"; SGP_B200_SO=sgp_b200/variants/libsgp_b200_$v.so timeout 120 python tools/profile_rt16.py 16 2>&1 | tail -1; done
echo "== stages 5"; timeout 120 python tools/profile_rt16.py 16 2>&1 | tail -1
echo "== ncu full tc16"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:reservoir_tc16 -s 2 -c 1 -o gpurun_out/r2_prof_scan16 python tools/profile_rt16.py 16 > gpurun_out/k_ncu_scan16.log 2>&1; tail -2 gpurun_out/k_ncu_scan16.log
