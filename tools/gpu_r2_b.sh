#!/bin/bash
# GPU call B: new tests (f-rows), REGDRAIN re-validation with tight timeouts, C5 with 4-step chunks.
mkdir -p gpurun_out
echo "== regdrain hop tests"; SGP_B200_SO=sgp_b200/variants/libsgp_b200_regdrain.so timeout 200 python -m pytest tests -m gpu -x -q -k "spmm_tensor_core" > gpurun_out/b_pytest_regdrain.log 2>&1; tail -3 gpurun_out/b_pytest_regdrain.log
echo "== regdrain timing"; SGP_B200_SO=sgp_b200/variants/libsgp_b200_regdrain.so timeout 150 python tools/profile_tc.py 16 > gpurun_out/b_tc_regdrain.txt 2>&1; cat gpurun_out/b_tc_regdrain.txt | tail -3
echo "== default timing"; timeout 150 python tools/profile_tc.py 16 > gpurun_out/b_tc_default.txt 2>&1; cat gpurun_out/b_tc_default.txt | tail -3
echo "== pytest -m gpu"; ( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/b_pytest.log 2>&1; tail -15 gpurun_out/b_pytest.log
echo "== bench c5"; timeout 600 python bench.py --steps 2 --warmup 3 --workload c5_1m --no-cpu > gpurun_out/b_bench_c5.json 2> gpurun_out/b_bench_c5.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/b_bench_c5.json').read().strip().splitlines()[-1])
print('c5 value %.1fM ms %.1f e2e %.1fM frac %.3f us/panel %.1f chunk %d'%(d['value']/1e6,d['ms_per_step'],d['e2e']['value']/1e6,d['roofline']['frac'],d['roofline']['us_per_hop_panel'],d['kernel_config']['chunk_steps']))
PY
