#!/bin/bash
# GPU call P (1 GPU): everything with the fp16x3 hop + scan as defaults.
mkdir -p gpurun_out
echo "== pytest -m gpu"; ( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/p_pytest.log 2>&1; grep -E "passed|failed|real" gpurun_out/p_pytest.log | tail -3
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
for W in c4_100k c5_1m c3_pv_us; do
  echo "== bench $W"; timeout 600 python bench.py --steps 3 --warmup 3 --workload $W > gpurun_out/p_bench_$W.json 2> gpurun_out/p_bench_$W.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/p_bench_$W.json").read().strip().splitlines()[-1])
    print("$W value %.1fM ms %.1f e2e %.1fM | hop %s frac %.3f (%.1f us/panel) | scan %.1f ms | cpu %.3fM | clocks %s | build %.0f ms"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6,d["roofline"]["kernel"][:22],d["roofline"]["frac"],d["roofline"]["us_per_hop_panel"],d["reservoir"]["ms_per_step"],d["cpu_baseline"]["value"]/1e6,d["clocks"]["sm_mhz"],d["e2e"]["operator_build_ms"]))
except Exception as e: print("$W unreadable", e); print(open("gpurun_out/p_bench_$W.err").read()[-1500:])
PY
done
