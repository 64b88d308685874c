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
CASES = [dict(N=3000, k=24, T=37, H=128, K=3, glob=True), dict(N=5016, k=100, T=20, H=256, K=2),
         dict(N=700, k=8, T=33, H=64, K=2, glob=True), dict(N=4000, k=90, T=21, H=128, K=2, bidir=True, glob=True),
         dict(N=3000, k=80, T=18, H=128, K=2, undirected=True, loops=True)]
for c in CASES:
    N, k, T, H, K = c["N"], c["k"], c["T"], c["H"], c["K"]
    ei, ew = sensor_knn(N, k, seed=0)
    x = sensor_signal(T, N, seed=1)
    torch.manual_seed(2)
    enc = sgp_b200.SGPEncoder(3, H, 1, 0.9, 0.9, 0.7, 1.0, K, c.get("bidir", False), False, c.get("glob", False),
                              add_self_loops=c.get("loops", False), undirected=c.get("undirected", False))
    enc.chunk_steps = 8
    full = enc(torch.from_numpy(x).to(dev), torch.from_numpy(ei).to(dev), torch.from_numpy(ew).to(dev))
    sh = RowShardedEncoder(enc, torch.from_numpy(ei), torch.from_numpy(ew), N, dev)
    out = torch.empty(T, sh.plan.n_own, enc.output_size, device=dev)
    acc = torch.zeros(1, dtype=torch.float64, device=dev)

    def sink(t0, t1, chunk):
        out[t0:t1].copy_(chunk)

    sh.encode_stream(torch.from_numpy(np.ascontiguousarray(x[:, sh.own])), sink, chunk_steps=5, checksum=acc)
    torch.cuda.synchronize()
    sh.check()
    ref = full[:, torch.from_numpy(sh.own).to(dev)]
    err = float((out - ref).abs().max() / ref.abs().max())
    cerr = abs(float(acc) - float(out.double().sum())) / max(abs(float(out.double().sum())), 1.0)
    worst = max(worst, err, cerr)
    print(f"rank {rank}: {c} own={sh.plan.n_own} halo={sh.halo_rows()} exchange={sh.exchange_used!r} "
          f"{getattr(sh, '_peer_error', '')} rel err {err:.2e} checksum err {cerr:.1e}", flush=True)
flag = torch.tensor([worst], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MAX)
if rank == 0:
    print("SHARDED CHECK", "OK" if float(flag) < 1e-5 else "FAILED", float(flag))
dist.destroy_process_group()
