#!/bin/bash
# GPU call H (1 GPU): scan kernel with CTA pairs sharing the W stream (TMA multicast) — parity + timing A/B.
mkdir -p gpurun_out
echo "== scan tests (pair)"; timeout 240 python -m pytest tests -m gpu -x -q -k "scan_tensor_core or baseline_shapes or fused_checksum or lockstep" > gpurun_out/h_pytest_pair.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/h_pytest_pair.log
echo "== timing pair";    timeout 120 python tools/profile_rt.py 16 2>&1 | tail -2
echo "== timing single";  SGP_B200_RT_PAIR=0 timeout 120 python tools/profile_rt.py 16 2>&1 | tail -2
echo "== bench c4 (pair)"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/h_bench_c4_pair.json 2> gpurun_out/h_bench_c4_pair.err
echo "== bench c4 (single)"; SGP_B200_RT_PAIR=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/h_bench_c4_single.json 2> gpurun_out/h_bench_c4_single.err
python - <<'PY'
import json
for n in ("pair","single"):
    try:
        d=json.loads(open("gpurun_out/h_bench_c4_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "value %.1fM ms %.1f e2e %.1fM scan %.1f ms hop frac %.3f (%.1f us/panel) clocks %s"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6,d["reservoir"]["ms_per_step"],d["roofline"]["frac"],d["roofline"]["us_per_hop_panel"],d["clocks"]["sm_mhz"]))
    except Exception as e: print(n, "unreadable", e)
PY
