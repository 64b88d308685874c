#!/bin/bash
# GPU call T (1 GPU): scan-under-hops experiment.
mkdir -p gpurun_out
timeout 420 python tools/overlap_experiment.py > gpurun_out/t_overlap.txt 2>&1
tail -20 gpurun_out/t_overlap.txt
