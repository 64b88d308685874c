#!/bin/bash
# GPU call G5 (N GPUs): C5 (N = 1M, 32-NN, H = 128, K = 2) row-sharded with the final kernels.
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 3 --warmup 2 --workload c5_1m > gpurun_out/g5_bench_n${N}_c5_1m.json 2> gpurun_out/g5_bench_n${N}_c5_1m.err
python - <<PY
import json
f="gpurun_out/g5_bench_n${N}_c5_1m.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value %.1fM ms %.1f e2e %.1fM halo %.3f frac %.3f build %.0f ms"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6,d["kernel_config"]["halo_rows_per_owned_row"],d["roofline"]["frac"],d["breakdown"]["operator_build_ms_max_over_ranks"]), {k:round(v,1) for k,v in d["breakdown"]["max_over_ranks"].items()}, d["clocks"], d["kernel_config"])
except Exception as e: print(f, "unreadable", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
