"""Per-step timestamps of one CTA of the tensor-core reservoir scan (SGP_B200_RT_TRACE).
Needs a library built with the stamps compiled in:
    python tools/build_variants.py rttrace=-DSGP_RT_TRACE_ON
    SGP_B200_SO=sgp_b200/variants/libsgp_b200_rttrace.so python tools/trace_rt.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sgp_b200 import ops
Tc = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N, H, Fin = 100000, 256, 1
dev = torch.device("cuda:0")
torch.manual_seed(0)
w_hh = (torch.rand(H, H, device=dev) * 2 - 1) * 0.05
w_ih = torch.rand(H, Fin, device=dev) * 2 - 1
b = torch.rand(H, device=dev) * 2 - 1
wimg = ops.reservoir_tc_pack(w_hh)
x = torch.randn(Tc, N, Fin, device=dev)
h = torch.zeros(N, H, device=dev)
out = torch.empty(Tc, N, H, device=dev)
err = torch.zeros(1, dtype=torch.int32, device=dev)
for _ in range(2):
    ops.reservoir_scan_tc(x, wimg, w_ih, b, 0.9, "tanh", h, out, err)
trace = torch.zeros(12 * 64, dtype=torch.int64, device=dev)
os.environ["SGP_B200_RT_TRACE"] = str(trace.data_ptr())
ops.reservoir_scan_tc(x, wimg, w_ih, b, 0.9, "tanh", h, out, err)
torch.cuda.synchronize()
del os.environ["SGP_B200_RT_TRACE"]
t = trace.view(12, 64).cpu()
base = int(t[0][0])
names = ["mma:start", "mma:h0 issued", "mma:h1 issued", "mma:W wait", "mma:state wait", "-",
         "epi:acc0", "epi:h0 computed", "epi:h0 published", "epi:acc1", "epi:h1 computed", "epi:h1 published"]
for step in range(min(Tc, 12)):
    print(f"step {step:2d} " + "  ".join(
        f"{names[r]}:{int(t[r][step]) - (0 if r in (3, 4) else base):7d}" for r in range(12) if r != 5))
d = (t[0][1:Tc] - t[0][:Tc - 1]).float()
print(f"cycles per step: mean {float(d.mean()):.0f} min {float(d.min()):.0f} max {float(d.max()):.0f}")
s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
s0.record()
for _ in range(3):
    ops.reservoir_scan_tc(x, wimg, w_ih, b, 0.9, "tanh", h, out, err)
s1.record(); torch.cuda.synchronize()
print(f"ms per time step {s0.elapsed_time(s1) / 3 / Tc:.4f}")
