#!/bin/bash
# GPU call X (1 GPU): gather-ring depth of the fp16x3 hop, A/B on one box (variants built by tools/build_variants.py).
mkdir -p gpurun_out
for rep in 1 2; do
  for v in default s10b3 s11b3 s12b2 s8b2; do
    if [ $v = default ]; then so=""; else so="sgp_b200/variants/libsgp_b200_$v.so"; fi
    echo "$v: $(SGP_B200_SO=$so timeout 300 python tools/profile_tc16.py 16 2>&1 | tail -2 | tr '\n' ' ')"
  done
done
