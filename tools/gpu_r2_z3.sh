#!/bin/bash
# GPU call Z3 (1 GPU): scan with the tanh constants folded into the epilogue FMAs — the whole GPU suite, smoke, A/B, bench.
mkdir -p gpurun_out
echo "== pytest -m gpu"; ( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/z3_pytest.log 2>&1; grep -E "passed|failed|real" gpurun_out/z3_pytest.log | tail -3
grep -E "^E " gpurun_out/z3_pytest.log | head -8
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
for rep in 1 2 3; do
  echo "folded: $(timeout 300 python tools/profile_rt16.py 16 2>&1 | tail -1)"
  echo "head:   $(SGP_B200_SO=sgp_b200/variants/libsgp_b200_head.so timeout 300 python tools/profile_rt16.py 16 2>&1 | tail -1)"
done
timeout 600 python bench.py --no-cpu > gpurun_out/z3_bench_c4.json 2> gpurun_out/z3_bench_c4.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/z3_bench_c4.json").read().strip().splitlines()[-1]); r = d["roofline"]
print("value %.1fM ms %.1f e2e %.1fM | frac %.3f (%.1f us/panel) | scan %.1f ms | clocks %s" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, r["frac"], r["us_per_hop_panel"], d["reservoir"]["ms_per_step"], d["clocks"]))
PY
