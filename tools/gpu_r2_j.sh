#!/bin/bash
# GPU call J (1 GPU): full suite with the fp16x3 scan as the default tanh path, smoke, bench A/B.
mkdir -p gpurun_out
echo "== pytest -m gpu"; ( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/j_pytest.log 2>&1; tail -6 gpurun_out/j_pytest.log | cut -c1-300
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== bench c4 (auto: fp16x3 scan)"; timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/j_bench_c4.json 2> gpurun_out/j_bench_c4.err
echo "== bench c4 (SGP_B200_RESERVOIR=tc: 3xTF32 scan)"; SGP_B200_RESERVOIR=tc timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/j_bench_c4_tf32.json 2> gpurun_out/j_bench_c4_tf32.err
python - <<'PY'
import json
for n in ("c4","c4_tf32"):
    try:
        d=json.loads(open("gpurun_out/j_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "value %.1fM ms %.1f e2e %.1fM scan %s %.1f ms hop frac %.3f (%.1f us/panel) clocks %s checksum %.6e"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6,d["reservoir"]["kernel"][:24],d["reservoir"]["ms_per_step"],d["roofline"]["frac"],d["roofline"]["us_per_hop_panel"],d["clocks"]["sm_mhz"],d["checksum"]))
    except Exception as e: print(n, "unreadable", e)
PY
