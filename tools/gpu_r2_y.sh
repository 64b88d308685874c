#!/bin/bash
# GPU call Y (1 GPU): final tree — the driver's three steps, then the ncu evidence refreshed for the final hop.
mkdir -p gpurun_out
echo "== pytest -m gpu"; ( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/y_pytest.log 2>&1; grep -E "passed|failed|real" gpurun_out/y_pytest.log | tail -3
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== bench (defaults)"; timeout 600 python bench.py > gpurun_out/y_bench_c4.json 2> gpurun_out/y_bench_c4.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/y_bench_c4.json").read().strip().splitlines()[-1]); r = d["roofline"]
print("value %.1fM ms %.1f e2e %.1fM | frac %.3f (%.1f us/panel) traffic %s | scan %.1f ms | cpu %.3fM | clocks %s | launches %s" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, r["frac"], r["us_per_hop_panel"], r["traffic"], d["reservoir"]["ms_per_step"], d["cpu_baseline"]["value"] / 1e6, d["clocks"], d["gpu_launches"]))
PY
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/y_bench_under_ncu.log 2>&1; python tools/launch_summary.py gpurun_out/r2_launches.csv > gpurun_out/r2_launches_summary.txt 2>&1; head -6 gpurun_out/r2_launches_summary.txt
echo "== ncu full hop16"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_rbu_tc16 -s 2 -c 1 -o gpurun_out/r2_prof_hop16 python tools/profile_tc16.py 16 > gpurun_out/y_ncu_hop16.log 2>&1; tail -2 gpurun_out/y_ncu_hop16.log
