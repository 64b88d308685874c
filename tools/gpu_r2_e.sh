#!/bin/bash
# GPU call E (N GPUs): hop chains in flight (lanes) A/B on the sharded bench, optional workload.
N=${1:-2}; W=${2:-c4_100k}; LANES=${3:-"2 3 4"}
mkdir -p gpurun_out
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
port=29520
for L in $LANES; do
  port=$((port+1))
  echo "== bench x$N $W lanes=$L"; SGP_B200_LANES=$L run 600 $port bench.py --gpus $N --steps 3 --warmup 2 --workload $W > gpurun_out/e_bench_n${N}_${W}_l$L.json 2> gpurun_out/e_bench_n${N}_${W}_l$L.err
  python - <<PY
import json
f="gpurun_out/e_bench_n${N}_${W}_l$L.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value %.1fM ms %.1f e2e %.1fM halo %.3f frac %.3f build %.0f ms"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6,d["kernel_config"]["halo_rows_per_owned_row"],d["roofline"]["frac"],d["breakdown"]["operator_build_ms_max_over_ranks"]), d["kernel_config"]["exchange"][:34], {k:round(v,1) for k,v in d["breakdown"]["max_over_ranks"].items()}, d["clocks"])
except Exception as e: print(f, "unreadable", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
done
