#!/bin/bash
# GPU call N (8 GPUs): SMs kept free of hop CTAs for the halo push — A/B on the C4 sharded bench (fp16x3 scan).
N=${1:-8}
mkdir -p gpurun_out
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
port=29540
for FREE in ${FREES:-0 12 24}; do
  port=$((port+1))
  echo "== bench x$N free_sms=$FREE"; SGP_B200_FREE_SMS=$FREE run 400 $port bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/n_bench_n${N}_free$FREE.json 2> gpurun_out/n_bench_n${N}_free$FREE.err
  python - <<PY
import json
f="gpurun_out/n_bench_n${N}_free$FREE.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value %.1fM ms %.1f e2e %.1fM"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6), {k:round(v,1) for k,v in d["breakdown"]["max_over_ranks"].items()}, d["clocks"]["sm_mhz"], d["checksum"])
except Exception as e: print(f, "unreadable", e); print(open(f.replace(".json",".err")).read()[-1200:])
PY
done
