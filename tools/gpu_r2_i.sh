#!/bin/bash
# GPU call I (1 GPU): bring-up of the fp16x3 tensor-core scan.
mkdir -p gpurun_out
echo "== fp16x3 scan tests (default packing)"; timeout 200 python -m pytest tests -m gpu -x -q -k "fp16x3" > gpurun_out/i_pytest.log 2>&1; echo "rc=$?"; grep -E "passed|failed|AssertionError|Error" gpurun_out/i_pytest.log | head -8
echo "== fp16x3 scan tests (swapped TMEM packing)"; SGP_B200_SO=sgp_b200/variants/libsgp_b200_r16swap.so timeout 200 python -m pytest tests -m gpu -x -q -k "fp16x3" > gpurun_out/i_pytest_swap.log 2>&1; echo "rc=$?"; grep -E "passed|failed|AssertionError|Error" gpurun_out/i_pytest_swap.log | head -8
echo "== timing fp16x3"; timeout 120 python tools/profile_rt16.py 16 2>&1 | tail -2
echo "== timing tf32";   timeout 120 python tools/profile_rt.py 16 2>&1 | tail -2
