#!/bin/bash
# GPU call Z5 (1 GPU): the whole GPU suite + smoke on the final library (one MMA issuer warp in the fp16x3 hop).
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/z5_pytest.log 2>&1; grep -E "passed|failed|real" gpurun_out/z5_pytest.log | tail -3
grep -E "^E " gpurun_out/z5_pytest.log | head -8
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
