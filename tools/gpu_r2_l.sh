#!/bin/bash
# GPU call L (1 GPU): fp16x3 scan after the epilogue rework — parity, timing, bench.
mkdir -p gpurun_out
echo "== fp16x3 scan tests"; timeout 200 python -m pytest tests -m gpu -x -q -k "fp16x3 or baseline_shapes" > gpurun_out/l_pytest.log 2>&1; echo "rc=$?"; grep -E "passed|failed|AssertionError|Error" gpurun_out/l_pytest.log | head -5
echo "== timing fp16x3"; timeout 120 python tools/profile_rt16.py 16 2>&1 | tail -1
echo "== bench c4"; timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/l_bench_c4.json 2> gpurun_out/l_bench_c4.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/l_bench_c4.json").read().strip().splitlines()[-1])
print("c4 value %.1fM ms %.1f e2e %.1fM scan %s %.1f ms hop frac %.3f (%.1f us/panel) clocks %s cpu %.3fM"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6,d["reservoir"]["kernel"][:24],d["reservoir"]["ms_per_step"],d["roofline"]["frac"],d["roofline"]["us_per_hop_panel"],d["clocks"]["sm_mhz"],d["cpu_baseline"]["value"]/1e6))
PY
