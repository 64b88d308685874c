"""Single-GPU experiment (C4): does running the next chunk's reservoir scan on a few SMs UNDER the
current chunk's hops beat running them back to back?  The hop is bound by the chip-wide L2 -> SM gather
rate rather than by SM count, so giving up SMs should cost it less than proportionally.
  1. hop alone at several persistent-CTA limits (sgp_tc_set_cta_limit);
  2. whole pass, sequential (the bench's schedule) against pipelined: hops limited to `limit` CTAs on one
     stream, the scan cut into launches of (148 - limit) node tiles on another.
python tools/overlap_experiment.py   (GPU box only)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sgp_b200 import ops, synthetic

cfg = synthetic.CONFIGS["c4_100k"]
N, T, H, K = cfg["N"], cfg["T"], cfg["H"], cfg["K"]
dev = torch.device("cuda:0")
enc = bench.make_encoder(cfg)
ei, ew, x = bench.make_inputs(cfg)
x = torch.from_numpy(x).to(dev)
F = H
D = (K + 1) * F
fwd, _ = enc.sgp_encoder.build_operators(torch.from_numpy(ei), torch.from_numpy(ew), N, dev, F)
plan = enc.reservoir.device_plan(dev, N)
acc = torch.zeros(1, dtype=torch.float64, device=dev)
Tc = 16
bufs = [torch.empty(Tc, N, D, device=dev) for _ in range(3)]
state = torch.zeros(1, N, H, device=dev)
bound0 = enc.reservoir.state_bound()


def hops(buf):
    b = bound0
    for h in range(1, K + 1):
        fwd.apply(buf[..., (h - 1) * F:h * F], buf[..., h * F:(h + 1) * F], checksum=acc, bound=b)
        b = fwd.out_bound(b)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3


# 1. hop alone
enc.reservoir.scan_chunk(plan, x[:Tc], state, bufs[0], acc)
for limit in (148, 136, 124, 112, 100):
    ops.tc_set_cta_limit(limit)
    ms = timed(lambda: hops(bufs[0]))
    print("hop alone, %3d CTAs: %.1f us per hop-panel" % (limit, ms * 1e3 / (K * Tc)), flush=True)
ops.tc_set_cta_limit(148)


def sequential():
    state.zero_()
    for i, t0 in enumerate(range(0, T, Tc)):
        t1 = min(T, t0 + Tc)
        buf = bufs[i % 3][: t1 - t0]
        enc.reservoir.scan_chunk(plan, x[t0:t1], state, buf, acc)
        hops(buf)


s_scan, s_hop = torch.cuda.Stream(), torch.cuda.Stream()


def pipelined(scan_tiles):
    state.zero_()
    main = torch.cuda.current_stream()
    s_scan.wait_stream(main)
    s_hop.wait_stream(main)
    ev_hop = []
    step_n = scan_tiles * 128 if scan_tiles else N
    for i, t0 in enumerate(range(0, T, Tc)):
        t1 = min(T, t0 + Tc)
        buf = bufs[i % 3][: t1 - t0]
        with torch.cuda.stream(s_scan):
            if i >= 3:
                s_scan.wait_event(ev_hop[i - 3])
            for n0 in range(0, N, step_n):
                n1 = min(N, n0 + step_n)
                enc.reservoir.scan_chunk(plan, x[t0:t1, n0:n1], state[:, n0:n1], buf[:, n0:n1], acc)
            ev = torch.cuda.Event()
            ev.record(s_scan)
        with torch.cuda.stream(s_hop):
            s_hop.wait_event(ev)
            hops(buf)
            e2 = torch.cuda.Event()
            e2.record(s_hop)
            ev_hop.append(e2)
    main.wait_stream(s_scan)
    main.wait_stream(s_hop)


acc.zero_(); sequential(); torch.cuda.synchronize(); ref = float(acc.item())
print("sequential: %.1f ms per pass" % timed(sequential), flush=True)
for limit, tiles in ((148, 0), (124, 24), (128, 20), (132, 16), (136, 12), (124, 0), (116, 32)):
    ops.tc_set_cta_limit(limit)
    acc.zero_(); pipelined(tiles); torch.cuda.synchronize(); got = float(acc.item())
    ms = timed(lambda: pipelined(tiles))
    print("pipelined, hop on %3d CTAs, scan in launches of %2d tiles: %.1f ms per pass (checksum rel diff %.1e)" % (
        limit, tiles, ms, abs(got - ref) / abs(ref)), flush=True)
ops.tc_set_cta_limit(148)
fwd.check(); enc.reservoir.check_plan(plan)
