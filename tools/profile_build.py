"""Where the per-graph operator build spends its time (C4 by default): wall-clock per stage with a
synchronize on both sides, then a torch.profiler table of one whole build.  GPU box only."""
import sys, time
import numpy as np
import torch
from sgp_b200 import synthetic, ops
from sgp_b200.preprocessing import build_operator
from sgp_b200.encoders import SGPEncoder

wl = sys.argv[1] if len(sys.argv) > 1 else "c4_100k"
cfg = synthetic.CONFIGS[wl]
N = cfg["N"]
ei, ew = synthetic.make_graph(cfg)
ei_h, ew_h = torch.from_numpy(ei), torch.from_numpy(ew)
dev = torch.device("cuda", 0)


def stamp(label, fn):
    torch.cuda.synchronize()
    t = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    print("%-28s %8.1f ms" % (label, (time.perf_counter() - t) * 1e3), flush=True)
    return r


for rep in range(2):
    print("--- pass", rep)
    op = stamp("build_operator (H2D + CSR)", lambda: build_operator(ei_h, ew_h, N, device=dev))
    csr = op.csr
    host = stamp("CSR D2H", lambda: (csr.rowptr.cpu().numpy(), csr.col.cpu().numpy(), csr.val.cpu().numpy()))
    for R in (64, 96):
        g = stamp("group_rows_host R=%d" % R, lambda: ops.group_rows_host(*host, N, R))
    stamp("tc16_build (given groups)", lambda: ops.tc16_build(csr, g))
    g64 = ops.group_rows_host(*host, N, 64)
    stamp("tc_build (given groups)", lambda: ops.tc_build(csr, g64))
    stamp("tc16_build (whole)", lambda: ops.tc16_build(csr))

from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    ops.tc16_build(csr, g)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=12, max_name_column_width=60))
