#!/bin/bash
# GPU call C (N GPUs): sharded correctness (P2P push exchange, then the NCCL fallback) + the N-GPU bench lines.
N=${1:-2}
mkdir -p gpurun_out
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
echo "== check_sharded x$N (auto: p2p)"; run 300 29511 tools/check_sharded.py > gpurun_out/c_check_sharded_n$N.log 2>&1; grep -E "SHARDED CHECK|exchange|Error|error" gpurun_out/c_check_sharded_n$N.log | tail -8
echo "== bench x$N (auto: p2p)"; run 400 29512 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/c_bench_n$N.json 2> gpurun_out/c_bench_n$N.err; tail -c 2600 gpurun_out/c_bench_n$N.json; grep -vE "OMP_NUM|^\*|^$" gpurun_out/c_bench_n$N.err | tail -5
echo "== bench x$N (nccl)"; SGP_B200_EXCHANGE=nccl run 400 29513 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/c_bench_n${N}_nccl.json 2> gpurun_out/c_bench_n${N}_nccl.err; python - <<PY
import json
for f in ("gpurun_out/c_bench_n$N.json","gpurun_out/c_bench_n${N}_nccl.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.1fM ms %.1f e2e %.1fM"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6), d["kernel_config"]["exchange"][:40], {k:round(v,1) for k,v in d["breakdown"]["max_over_ranks"].items()})
    except Exception as e: print(f, "unreadable", e)
PY
