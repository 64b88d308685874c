#!/bin/bash
# GPU call V (1 GPU): balanced work order of the fp16x3 hop — parity, A/B standalone, bench line.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q -k "fp16x3 or tc16 or sharded or c4 or c5 or checksum" ) > gpurun_out/v_pytest.log 2>&1; grep -E "passed|failed" gpurun_out/v_pytest.log | tail -2
grep -E "^E |Error" gpurun_out/v_pytest.log | head -10
echo "balanced:"; timeout 300 python tools/profile_tc16.py 16 2>&1 | tail -1
echo "plain:"; SGP_B200_TC16_BALANCE=0 timeout 300 python tools/profile_tc16.py 16 2>&1 | tail -1
timeout 600 python bench.py --no-cpu > gpurun_out/v_bench_c4.json 2> gpurun_out/v_bench_c4.err
timeout 600 python bench.py --no-cpu --workload c5_1m > gpurun_out/v_bench_c5.json 2> gpurun_out/v_bench_c5.err
python - <<'PY'
import json
for f in ("v_bench_c4", "v_bench_c5"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1]); r = d["roofline"]
        print("%s value %.1fM ms %.1f e2e %.1fM | frac %.3f (%.1f us/panel) | scan %.1f ms | clocks %s" % (f, d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, r["frac"], r["us_per_hop_panel"], d["reservoir"]["ms_per_step"], d["clocks"]))
    except Exception as e:
        print(f, "unreadable", e); print(open("gpurun_out/%s.err" % f).read()[-1500:])
PY
