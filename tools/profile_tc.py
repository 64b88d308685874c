"""Driver for the tensor-core hop at C4 shapes: python tools/profile_tc.py [Tc] [--sweep]
Times the hop with CUDA events (buffers sized like the encoder's chunk: row stride 1280 floats)."""
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
tc = ops.tc_build(op.csr)
buf = torch.randn(Tc, N, 5 * H, device=dev)
acc = torch.zeros(1, dtype=torch.float64, device=dev)      # the bench's sink: checksum fused into the epilogue


def run(reps=3):
    for _ in range(2):
        ops.spmm_tc(tc, buf[..., :H], buf[..., H:2 * H], checksum=acc)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.spmm_tc(tc, buf[..., :H], buf[..., H:2 * H], checksum=acc)
    e1.record(); torch.cuda.synchronize(); ops.tc_check(tc)
    return e0.elapsed_time(e1) / reps / Tc * 1e3


if "--sweep" in sys.argv:
    for pw in (2, 4, 8):
        os.environ["SGP_B200_TC_PW"] = str(pw)
        print(f"producer warps {pw}: {run():.1f} us per hop-panel")
else:
    print(f"{run():.1f} us per hop-panel")
