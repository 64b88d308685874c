#!/bin/bash
# GPU call D: REGDRAIN at 88 registers (tight timeouts), then the round-2 ncu evidence.
mkdir -p gpurun_out
echo "== regdrain hop tests"; SGP_B200_SO=sgp_b200/variants/libsgp_b200_regdrain.so timeout 100 python -m pytest tests -m gpu -x -q -k "spmm_tensor_core" > gpurun_out/d_pytest_regdrain.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/d_pytest_regdrain.log
echo "== regdrain timing"; SGP_B200_SO=sgp_b200/variants/libsgp_b200_regdrain.so timeout 100 python tools/profile_tc.py 16 > gpurun_out/d_tc_regdrain.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/d_tc_regdrain.txt
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/d_bench_under_ncu.log 2>&1; python tools/launch_summary.py gpurun_out/r2_launches.csv | tee gpurun_out/r2_launches_summary.txt | head -20
echo "== ncu full hop"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_rbu_tc -s 2 -c 1 -o gpurun_out/r2_prof_hop python tools/profile_tc.py 16 > gpurun_out/d_ncu_hop.log 2>&1; tail -2 gpurun_out/d_ncu_hop.log
echo "== ncu full scan"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:reservoir_tc -s 2 -c 1 -o gpurun_out/r2_prof_scan python tools/profile_rt.py 16 > gpurun_out/d_ncu_scan.log 2>&1; tail -2 gpurun_out/d_ncu_scan.log
ls -la gpurun_out/*.ncu-rep | tail -3
