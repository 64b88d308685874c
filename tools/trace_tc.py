"""Per-item timestamps of one CTA of the tensor-core hop (SGP_B200_TC_TRACE): which role waits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sgp_b200 import ops
from sgp_b200.preprocessing import build_operator
from sgp_b200.synthetic import CONFIGS, make_graph
cfg = CONFIGS["c4_100k"]; Tc = 4
dev = torch.device("cuda:0"); N, H = cfg["N"], cfg["H"]
ei, ew = make_graph(cfg, seed=0)
op = build_operator(torch.from_numpy(ei), torch.from_numpy(ew), N, device=dev)
tc = ops.tc_build(op.csr)
buf = torch.randn(Tc, N, 2 * H, device=dev)
trace = torch.zeros(12 * 512, dtype=torch.int64, device=dev)
for _ in range(2):
    ops.spmm_tc(tc, buf[..., :H], buf[..., H:])
os.environ["SGP_B200_TC_TRACE"] = str(trace.data_ptr())
ops.spmm_tc(tc, buf[..., :H], buf[..., H:])
torch.cuda.synchronize()
t = trace.view(12, 512).cpu()
import numpy as np
os.makedirs("gpurun_out", exist_ok=True); np.save("gpurun_out/trace_tc.npy", t.numpy())
n = int((t[5] > 0).sum()); base = int(t[0][0])
names = ["prod:issue", "prod:full", "split:full", "split:afree", "split:ready", "mma:ready", "mma:issued", "mma:wait", "rdy q0", "rdy q1", "rdy q2", "rdy q3"]
print(f"items traced {n}; cycles relative to the producer's first issue")
for i in list(range(0, 6)) + list(range(20, 32)):
    print(f"item {i:3d} " + "  ".join(f"{names[r]}:{(int(t[r][i]) - base) if int(t[r][i]) else -1:7d}" for r in (0, 2, 3, 8, 9, 10, 11, 7, 5, 6)))
d = (t[6][1:n] - t[6][:n - 1]).float()
print(f"mean cycles between successive MMA issues: {float(d.mean()):.0f} (min {float(d.min()):.0f} max {float(d.max()):.0f})")
