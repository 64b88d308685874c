#!/bin/bash
# GPU call S (1 GPU): final tree — the driver's three steps, the operator-build breakdown, chunk-length sweep.
mkdir -p gpurun_out
echo "== pytest -m gpu"; ( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/s_pytest.log 2>&1; grep -E "passed|failed|real" gpurun_out/s_pytest.log | tail -3
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== bench (defaults)"; timeout 600 python bench.py > gpurun_out/s_bench_c4.json 2> gpurun_out/s_bench_c4.err
PYTHONPATH=. timeout 300 python tools/profile_build.py c4_100k > gpurun_out/s_profile_build_c4.txt 2>&1
grep -E " ms$|pass" gpurun_out/s_profile_build_c4.txt | tail -8
for C in 40 100; do
  timeout 600 python bench.py --no-cpu --chunk $C > gpurun_out/s_bench_c4_chunk$C.json 2> gpurun_out/s_bench_c4_chunk$C.err
done
python - <<'PY'
import json
for f in ("s_bench_c4", "s_bench_c4_chunk40", "s_bench_c4_chunk100"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print("%s value %.1fM ms %.1f e2e %.1fM | frac %.3f (%.1f us/panel) traffic %s | scan %.1f ms | clocks %s | build %.0f ms (first %.0f)" % (
            f, d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, r["frac"], r["us_per_hop_panel"],
            r["traffic"], d["reservoir"]["ms_per_step"], d["clocks"], d["e2e"]["operator_build_ms"],
            d["e2e"]["operator_build_first_call_ms"]))
    except Exception as e:
        print(f, "unreadable", e); print(open("gpurun_out/%s.err" % f).read()[-1500:])
PY
