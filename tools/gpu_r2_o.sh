#!/bin/bash
# GPU call O (1 GPU): bring-up of the fp16x3 / 96-row hop.
mkdir -p gpurun_out
echo "== fp16x3 hop tests"; timeout 240 python -m pytest tests -m gpu -x -q -k "spmm_fp16x3 or fp16x3_hop" > gpurun_out/o_pytest.log 2>&1; echo "rc=$?"; grep -E "passed|failed|AssertionError|Error|error" gpurun_out/o_pytest.log | head -8
echo "== timing fp16x3 hop"; timeout 200 python tools/profile_tc16.py 16 2>&1 | tail -2
echo "== timing tf32 hop"; timeout 200 python tools/profile_tc.py 16 2>&1 | tail -1
