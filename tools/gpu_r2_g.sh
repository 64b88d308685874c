#!/bin/bash
# GPU call G (N GPUs): sharded check + C4 bench (+ optional second workload) with the current exchange.
N=${1:-2}; W2=${2:-}
mkdir -p gpurun_out
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
echo "== check_sharded x$N"; run 300 29531 tools/check_sharded.py > gpurun_out/g_check_sharded_n$N.log 2>&1; grep -E "SHARDED CHECK|Error|error" gpurun_out/g_check_sharded_n$N.log | tail -5
for W in c4_100k $W2; do
  echo "== bench x$N $W"; run 900 $((29532 + ${#W})) bench.py --gpus $N --steps 3 --warmup 2 --workload $W > gpurun_out/g_bench_n${N}_$W.json 2> gpurun_out/g_bench_n${N}_$W.err
  python - <<PY
import json
f="gpurun_out/g_bench_n${N}_$W.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value %.1fM ms %.1f e2e %.1fM halo %.3f frac %.3f build %.0f ms"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6,d["kernel_config"]["halo_rows_per_owned_row"],d["roofline"]["frac"],d["breakdown"]["operator_build_ms_max_over_ranks"]), {k:round(v,1) for k,v in d["breakdown"]["max_over_ranks"].items()}, d["clocks"])
except Exception as e: print(f, "unreadable", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
done
