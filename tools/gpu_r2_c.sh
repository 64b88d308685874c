#!/bin/bash
# GPU call C (N GPUs): sharded correctness under NCCL + the N-GPU bench line.
N=${1:-2}
mkdir -p gpurun_out
echo "== check_sharded x$N"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py > gpurun_out/c_check_sharded_n$N.log 2>&1; grep -E "SHARDED CHECK|rank 0|Error|error" gpurun_out/c_check_sharded_n$N.log | tail -12
echo "== bench x$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/c_bench_n$N.json 2> gpurun_out/c_bench_n$N.err; tail -c 3500 gpurun_out/c_bench_n$N.json; tail -5 gpurun_out/c_bench_n$N.err
