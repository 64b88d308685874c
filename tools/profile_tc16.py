"""Driver for the fp16x3 / 96-row tensor-core hop at C4 shapes: python tools/profile_tc16.py [Tc]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sgp_b200 import ops
from sgp_b200.preprocessing import build_operator
from sgp_b200.synthetic import CONFIGS, make_graph
cfg = CONFIGS["c4_100k"]; Tc = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4
dev = torch.device("cuda:0"); N, H = cfg["N"], cfg["H"]
ei, ew = make_graph(cfg, seed=0)
op = build_operator(torch.from_numpy(ei), torch.from_numpy(ew), N, device=dev)
tc = ops.tc16_build(op.csr)
print("groups", tc.n_groups, "chunks", int(tc.chunk_ptr[-1]), "fill %.3f" % tc.fill, "U/R %.2f" % (100 / tc.fill / 96 if tc.fill else 0))
buf = torch.tanh(torch.randn(Tc, N, 5 * H, device=dev))
acc = torch.zeros(1, dtype=torch.float64, device=dev)
for _ in range(2):
    ops.spmm_tc16(tc, buf[..., :H], buf[..., H:2 * H], 1.0, checksum=acc)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    ops.spmm_tc16(tc, buf[..., :H], buf[..., H:2 * H], 1.0, checksum=acc)
e1.record(); torch.cuda.synchronize(); ops.tc_check(tc)
print(f"{e0.elapsed_time(e1) / 3 / Tc * 1e3:.1f} us per hop-panel")
chk = torch.empty(Tc, N, H, device=dev)
ops.spmm(op.csr, buf[..., :H], chk)
print("max |tc16 - csr| / max |csr| = %.2e" % float((buf[..., H:2 * H] - chk).abs().max() / chk.abs().max()))
