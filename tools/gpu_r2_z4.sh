#!/bin/bash
# GPU call Z4 (1 GPU): MMA issuer warps of the fp16x3 hop: 2 (default) against 4 and 1 on one box; then the whole GPU suite on the 4-issuer library.
mkdir -p gpurun_out
for rep in 1 2; do
  for v in default iss4 iss1; do
    if [ $v = default ]; then so=""; else so="sgp_b200/variants/libsgp_b200_$v.so"; fi
    echo "$v: $(SGP_B200_SO=$so timeout 200 python tools/profile_tc16.py 16 2>&1 | tail -2 | tr '\n' ' ')"
  done
done
echo "== pytest -m gpu (4 issuers)"; ( time SGP_B200_SO=sgp_b200/variants/libsgp_b200_iss4.so timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/z4_pytest.log 2>&1; grep -E "passed|failed|real" gpurun_out/z4_pytest.log | tail -3
grep -E "^E " gpurun_out/z4_pytest.log | head -8
