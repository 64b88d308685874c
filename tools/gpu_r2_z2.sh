#!/bin/bash
# GPU call Z2 (1 GPU): fp16x3 hop gathering with per-row 512-byte bulk copies against cp.async — parity subset, A/B on one box.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q -x -k "fp16x3 or tc16 or sharded or c4 or c5 or checksum" ) > gpurun_out/z2_pytest.log 2>&1; grep -E "passed|failed" gpurun_out/z2_pytest.log | tail -2
grep -E "^E " gpurun_out/z2_pytest.log | head -8
for rep in 1 2 3; do
  echo "bulk rows: $(timeout 300 python tools/profile_tc16.py 16 2>&1 | tail -2 | tr '\n' ' ')"
  echo "cp.async:  $(SGP_B200_SO=sgp_b200/variants/libsgp_b200_ldgsts.so timeout 300 python tools/profile_tc16.py 16 2>&1 | tail -2 | tr '\n' ' ')"
done
