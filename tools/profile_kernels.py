"""Small driver for ncu: launches each hot kernel a few times at the C4 shapes
(N=100k, k=100-NN, H=F=256) with a short time chunk so that `ncu --set full` replays stay cheap.

    ncu --set full --clock-control none --import-source on -k regex:'spmm_rbu|reservoir_scan' \
        -o gpurun_out/prof python tools/profile_kernels.py --tc 2
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import sgp_b200  # noqa: E402
from sgp_b200 import ops  # noqa: E402
from sgp_b200.preprocessing import build_operator  # noqa: E402
from sgp_b200.synthetic import CONFIGS, make_graph, sensor_signal  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c4_100k")
ap.add_argument("--tc", type=int, default=2)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--rbu", default="force16")
ap.add_argument("--skip-csr", action="store_true")
args = ap.parse_args()

cfg = CONFIGS[args.workload]
dev = torch.device("cuda:0")
N, H, Fin = cfg["N"], cfg["H"], cfg["Fin"]
ei, ew = make_graph(cfg, seed=0)
op = build_operator(torch.from_numpy(ei), torch.from_numpy(ew), N, device=dev)
op.maybe_build_rbu(H, args.rbu)
print("nnz", op.csr.nnz, "rbu", None if op.rbu is None else (op.rbu.R, round(op.rbu.fill, 3), int(op.rbu.ucol.numel())))
x = torch.from_numpy(sensor_signal(args.tc, N, seed=1, exogenous=Fin == 3)).to(dev)
torch.manual_seed(2)
res = sgp_b200.Reservoir(Fin, H, density=0.7)
plan = res.device_plan(dev, N)
print('reservoir path:', plan[0][0])
buf = torch.zeros(args.tc, N, 3 * H, device=dev)
state = torch.zeros(1, N, H, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
for _ in range(args.reps):
    ev[0].record()
    res.scan_chunk(plan, x, state, buf)
    ev[1].record()
    op.apply(buf[..., :H], buf[..., H:2 * H])
    ev[2].record()
    if not args.skip_csr:
        ops.spmm(op.csr, buf[..., :H], buf[..., 2 * H:])
    ev[3].record()
torch.cuda.synchronize()
print("scan ms/step %.3f  rbu ms/panel %.3f  csr ms/panel %.3f" % (
    ev[0].elapsed_time(ev[1]) / args.tc, ev[1].elapsed_time(ev[2]) / args.tc,
    ev[2].elapsed_time(ev[3]) / args.tc))
if not args.skip_csr:
    print("max |rbu - csr| =", float((buf[..., H:2 * H] - buf[..., 2 * H:]).abs().max()))
