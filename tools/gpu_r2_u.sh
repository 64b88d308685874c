#!/bin/bash
# GPU call U (1 GPU): hop variant check — parity tests of the fp16x3 hop, standalone timing, bench line.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q -k "fp16x3 or tc16 or sharded or c4 or c5 or checksum" ) > gpurun_out/u_pytest.log 2>&1; grep -E "passed|failed" gpurun_out/u_pytest.log | tail -2
timeout 300 python tools/profile_tc16.py 16 2>&1 | tail -2
timeout 600 python bench.py --no-cpu > gpurun_out/u_bench_c4.json 2> gpurun_out/u_bench_c4.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/u_bench_c4.json").read().strip().splitlines()[-1]); r = d["roofline"]
print("value %.1fM ms %.1f e2e %.1fM | frac %.3f (%.1f us/panel) | scan %.1f ms | clocks %s" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, r["frac"], r["us_per_hop_panel"], d["reservoir"]["ms_per_step"], d["clocks"]))
PY
