"""Sweep the RBU SpMM launch variants on the C4 graph (device timing, CUDA events)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sgp_b200 import ops  # noqa: E402
from sgp_b200.preprocessing import build_operator  # noqa: E402
from sgp_b200.synthetic import CONFIGS, make_graph  # noqa: E402

cfg = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c4_100k"]
Tc = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda:0")
N, H = cfg["N"], cfg["H"]
ei, ew = make_graph(cfg, seed=0)
op = build_operator(torch.from_numpy(ei), torch.from_numpy(ew), N, device=dev)
rbus = {R: ops.rbu_build(op.csr, R) for R in (8, 16)}
for R, r in rbus.items():
    print(f"R={R} fill={r.fill:.3f} U/R={r.ucol.numel() / N:.2f}")
buf = torch.randn(Tc, N, 3 * H, device=dev)
src, dst, ref = buf[..., :H], buf[..., H:2 * H], buf[..., 2 * H:]
ops.spmm(op.csr, src, ref)
flops = 2 * op.csr.nnz * H


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps / Tc * 1e3   # us per hop-panel


print("csr            %8.1f us/panel" % timeit(lambda: ops.spmm(op.csr, src, dst)))
tc = ops.tc_build(op.csr)
us = timeit(lambda: ops.spmm_tc(tc, src, dst))
ops.tc_check(tc)
print(f"tcgen05 R=64 fill={tc.fill:.3f}  {us:8.1f} us/panel  {flops / us / 1e6:6.2f} TF/s useful  "
      f"alg {(8 * op.csr.nnz + 8 * N * H) / us / 1e3:7.1f} GB/s  err={float((dst - ref).abs().max() / ref.abs().max()):.2e}")
for R in (16,):
    for ver, minb, tspan in [(1, 4, 0), (2, 4, 2), (2, 3, 2), (3, 4, 2), (3, 4, 4)]:
        if R == 8 and minb == 3:
            continue
        os.environ.update(SGP_B200_RBU_KERNEL=str(ver), SGP_B200_RBU_MINB=str(minb), SGP_B200_RBU_TSPAN=str(tspan))
        us = timeit(lambda: ops.spmm_rbu(rbus[R], src, dst))
        err = float((dst - ref).abs().max())
        print(f"R={R:2d} v{ver} minb={minb} tspan={tspan:2d}  {us:8.1f} us/panel  {flops / us / 1e6:6.2f} TF/s useful  "
              f"{flops / rbus[R].fill / us / 1e6:6.2f} TF/s issued  err={err:.2e}")
