"""ncu driver for the tensor-core hop at C4 shapes: python tools/profile_tc.py [Tc]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sgp_b200 import ops
from sgp_b200.preprocessing import build_operator
from sgp_b200.synthetic import CONFIGS, make_graph
cfg = CONFIGS["c4_100k"]; Tc = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0"); N, H = cfg["N"], cfg["H"]
ei, ew = make_graph(cfg, seed=0)
op = build_operator(torch.from_numpy(ei), torch.from_numpy(ew), N, device=dev)
tc = ops.tc_build(op.csr)
buf = torch.randn(Tc, N, 2 * H, device=dev)
for _ in range(3):
    ops.spmm_tc(tc, buf[..., :H], buf[..., H:])
torch.cuda.synchronize(); ops.tc_check(tc); print("ok")
