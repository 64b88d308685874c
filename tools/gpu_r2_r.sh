#!/bin/bash
# GPU call R (1 GPU): operator-build breakdown.
mkdir -p gpurun_out
PYTHONPATH=. timeout 600 python tools/profile_build.py c4_100k > gpurun_out/r_profile_build_c4.txt 2>&1
grep -E " ms$|pass" gpurun_out/r_profile_build_c4.txt | head -40
