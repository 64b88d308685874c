#!/bin/bash
# GPU call W (1 GPU): A/B on ONE box — fp16x3 hop with the conversion under the MMAs (default) against the earlier order.
mkdir -p gpurun_out
for rep in 1 2 3; do
  echo "under-mma:"; timeout 300 python tools/profile_tc16.py 16 2>&1 | tail -1
  echo "late:";      SGP_B200_SO=sgp_b200/variants/libsgp_b200_late.so timeout 300 python tools/profile_tc16.py 16 2>&1 | tail -1
done
