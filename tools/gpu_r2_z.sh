#!/bin/bash
# GPU call Z (1 GPU): fp16x3 hop with the A tile handed over per k-step — parity subset, A/B against HEAD's library on one box.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q -x -k "fp16x3 or tc16 or sharded or c4 or c5 or checksum" ) > gpurun_out/z_pytest.log 2>&1; grep -E "passed|failed" gpurun_out/z_pytest.log | tail -2
grep -E "^E " gpurun_out/z_pytest.log | head -8
for rep in 1 2 3; do
  echo "k-step handover: $(timeout 300 python tools/profile_tc16.py 16 2>&1 | tail -2 | tr '\n' ' ')"
  echo "head:            $(SGP_B200_SO=sgp_b200/variants/libsgp_b200_head.so timeout 300 python tools/profile_tc16.py 16 2>&1 | tail -2 | tr '\n' ' ')"
done
