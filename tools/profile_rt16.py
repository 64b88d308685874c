"""Timing driver for the fp16x3 tensor-core scan at C4 shapes (one launch = Tc steps of all N nodes, writing
into a [Tc, N, 5H] chunk like the encoder does):  python tools/profile_rt16.py [Tc]"""
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
wimg, scale = ops.reservoir_tc16_pack(w_hh)
x = torch.randn(Tc, N, Fin, device=dev)
h = torch.zeros(N, H, device=dev)
buf = torch.empty(Tc, N, 5 * H, device=dev)
err = torch.zeros(1, dtype=torch.int32, device=dev)
acc = torch.zeros(1, dtype=torch.float64, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2):
    ops.reservoir_scan_tc16(x, wimg, scale, w_ih, b, 0.9, h, buf[..., :H], err, acc)
e0.record()
for _ in range(3):
    ops.reservoir_scan_tc16(x, wimg, scale, w_ih, b, 0.9, h, buf[..., :H], err, acc)
e1.record(); torch.cuda.synchronize()
assert int(err.item()) == 0
print(f"{e0.elapsed_time(e1) / 3 / Tc * 1e3:.1f} us per time step")
