#!/bin/bash
# GPU call M (1 GPU): refresh the round-2 evidence with the fp16x3 scan as the default path.
mkdir -p gpurun_out
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/m_bench_under_ncu.log 2>&1; python tools/launch_summary.py gpurun_out/r2_launches.csv > gpurun_out/r2_launches_summary.txt 2>&1; head -6 gpurun_out/r2_launches_summary.txt
echo "== ncu full tc16"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:reservoir_tc16 -s 2 -c 1 -o gpurun_out/r2_prof_scan16_v2 python tools/profile_rt16.py 16 > gpurun_out/m_ncu_scan16.log 2>&1; tail -1 gpurun_out/m_ncu_scan16.log
echo "== bench c5 / c3 (fp16x3 scan)"
for W in c5_1m c3_pv_us; do timeout 600 python bench.py --steps 2 --warmup 3 --workload $W > gpurun_out/m_bench_$W.json 2> gpurun_out/m_bench_$W.err; python - <<PY
import json
d=json.loads(open("gpurun_out/m_bench_$W.json").read().strip().splitlines()[-1])
print("$W value %.1fM ms %.1f e2e %.1fM scan %s %.1f ms hop frac %.3f cpu %.3fM"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6,d["reservoir"]["kernel"][:22],d["reservoir"]["ms_per_step"],d["roofline"]["frac"],d["cpu_baseline"]["value"]/1e6))
PY
done
