#!/bin/bash
# GPU call F (1 GPU): full parity suite with the new tests, smoke(), shipped-config bench lines.
mkdir -p gpurun_out
echo "== pytest -m gpu"; ( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/f_pytest.log 2>&1; tail -6 gpurun_out/f_pytest.log
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
for W in la_yaml pv_yaml; do
  echo "== bench $W"; timeout 600 python bench.py --steps 2 --warmup 3 --workload $W > gpurun_out/f_bench_$W.json 2> gpurun_out/f_bench_$W.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/f_bench_$W.json").read().strip().splitlines()[-1])
    print("$W value %.2fM ms %.1f e2e %.2fM scan %s %.1f ms cpu %.3fM (%d cores)"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6,d["reservoir"]["kernel"][:28],d["reservoir"]["ms_per_step"],d["cpu_baseline"]["value"]/1e6,d["cpu_baseline"]["cores"]), d["kernel_config"])
except Exception as e: print("$W unreadable", e); print(open("gpurun_out/f_bench_$W.err").read()[-1500:])
PY
done
echo "== bench $W layerwise (round-1 path: one launch per layer)"; SGP_B200_RESERVOIR=layerwise timeout 600 python bench.py --steps 2 --warmup 3 --workload pv_yaml --no-cpu > gpurun_out/f_bench_pv_yaml_layerwise.json 2> gpurun_out/f_bench_pv_yaml_layerwise.err
python - <<PY
import json
for W in ("pv_yaml_layerwise",):
    try:
        d=json.loads(open("gpurun_out/f_bench_%s.json"%W).read().strip().splitlines()[-1])
        print(W, "value %.2fM ms %.1f scan %s %.1f ms"%(d["value"]/1e6,d["ms_per_step"],d["reservoir"]["kernel"][:28],d["reservoir"]["ms_per_step"]))
    except Exception as e: print(W, "unreadable", e)
PY
