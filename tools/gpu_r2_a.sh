#!/bin/bash
# GPU call A of round 2: parity suite, REGDRAIN validation, L2 ingest microbenchmark, bench lines.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/a_smi.txt 2>&1
echo "== pytest -m gpu" ; ( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/a_pytest.log 2>&1; tail -5 gpurun_out/a_pytest.log
echo "== l2_ingest"; timeout 120 tools/microbench/l2_ingest > gpurun_out/a_l2_ingest.txt 2>&1; cat gpurun_out/a_l2_ingest.txt
echo "== hop timing default"; timeout 300 python tools/profile_tc.py 16 > gpurun_out/a_tc_default.txt 2>&1; cat gpurun_out/a_tc_default.txt
echo "== hop regdrain parity + timing"
SGP_B200_SO=sgp_b200/variants/libsgp_b200_regdrain.so timeout 600 python -m pytest tests -m gpu -x -q -k "tensor_core or baseline_shapes or lockstep or fused_checksum" > gpurun_out/a_pytest_regdrain.log 2>&1; tail -3 gpurun_out/a_pytest_regdrain.log
SGP_B200_SO=sgp_b200/variants/libsgp_b200_regdrain.so timeout 300 python tools/profile_tc.py 16 > gpurun_out/a_tc_regdrain.txt 2>&1; cat gpurun_out/a_tc_regdrain.txt
echo "== bench c4"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/a_bench_c4.json 2> gpurun_out/a_bench_c4.err; tail -c 3000 gpurun_out/a_bench_c4.json; tail -3 gpurun_out/a_bench_c4.err
echo "== bench c5"; timeout 900 python bench.py --steps 2 --warmup 3 --workload c5_1m > gpurun_out/a_bench_c5.json 2> gpurun_out/a_bench_c5.err; tail -c 2500 gpurun_out/a_bench_c5.json; tail -3 gpurun_out/a_bench_c5.err
echo "== bench c3"; timeout 900 python bench.py --steps 2 --warmup 3 --workload c3_pv_us > gpurun_out/a_bench_c3.json 2> gpurun_out/a_bench_c3.err; tail -c 1500 gpurun_out/a_bench_c3.json; tail -3 gpurun_out/a_bench_c3.err
echo "== bench c2"; timeout 900 python bench.py --steps 2 --warmup 3 --workload c2_pems_bay > gpurun_out/a_bench_c2.json 2> gpurun_out/a_bench_c2.err; tail -c 1500 gpurun_out/a_bench_c2.json; tail -3 gpurun_out/a_bench_c2.err
echo "== reference arm c4"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/a_ref_c4.json 2> gpurun_out/a_ref_c4.err; cat gpurun_out/a_ref_c4.json
