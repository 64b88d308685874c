"""torchrun --nproc-per-node N tools/check_sharded.py : the row-sharded encoder against the
single-GPU encoder (both CUDA) on every rank, small shapes; then a timing of the exchange."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import sgp_b200  # noqa: E402
from sgp_b200.sharded import RowShardedEncoder  # noqa: E402
from sgp_b200.synthetic import sensor_knn, sensor_signal  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
worst = 0.0
for (N, k, T, H, K, glob) in [(3000, 24, 37, 128, 3, True), (5016, 100, 20, 256, 2, False), (700, 8, 33, 64, 2, True)]:
    ei, ew = sensor_knn(N, k, seed=0)
    x = sensor_signal(T, N, seed=1)
    torch.manual_seed(2)
    enc = sgp_b200.SGPEncoder(3, H, 1, 0.9, 0.9, 0.7, 1.0, K, False, False, glob)
    enc.chunk_steps = 8
    full = enc(torch.from_numpy(x).to(dev), torch.from_numpy(ei).to(dev), torch.from_numpy(ew).to(dev))
    sh = RowShardedEncoder(enc, torch.from_numpy(ei), torch.from_numpy(ew), N, dev)
    out = torch.empty(T, sh.plan.n_own, enc.output_size, device=dev)

    def sink(t0, t1, chunk):
        out[t0:t1].copy_(chunk)

    sh.encode_stream(torch.from_numpy(np.ascontiguousarray(x[:, sh.own])), sink, chunk_steps=8)
    torch.cuda.synchronize()
    ref = full[:, torch.from_numpy(sh.own).to(dev)]
    err = float((out - ref).abs().max() / ref.abs().max())
    worst = max(worst, err)
    print(f"rank {rank}: N={N} own={sh.plan.n_own} halo={sh.plan.n_halo} rel err {err:.2e}", flush=True)
flag = torch.tensor([worst], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MAX)
if rank == 0:
    print("SHARDED CHECK", "OK" if float(flag) < 1e-5 else "FAILED", float(flag))
dist.destroy_process_group()
