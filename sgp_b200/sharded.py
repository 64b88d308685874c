"""Row-sharded encoder over the GPUs of one box (one process per GPU, torch.distributed / NCCL).

The reference is single-process; this is the B200-side scaling design named by the north star:

* the reservoir is embarrassingly parallel over nodes (shared frozen weights, per-node state) —
  every rank scans only its own rows, no communication;
* the K-hop propagation is sharded by DESTINATION rows: rank r owns a contiguous range of the
  locality-ordered row groups (the same breadth-first greedy groups the RBU kernel uses), i.e. a
  compact patch of the sensor graph, plus the matching rows of every feature block;
* before each hop the rows of the previous block that other ranks reference ("halo" rows) are
  packed on the device (sgp_gather_rows) and exchanged with ONE all-to-all-v over NVLink; the
  SpMM kernels then read local columns from the rank's own block and halo columns straight
  from the receive buffer (second source pointer, no concatenation copy);
* two chunks are in flight on two CUDA streams so that the exchange of one overlaps the SpMM of
  the other.

The partition / halo plan is plain numpy, identical on every rank (deterministic from the CSR),
and is what the world_size-2 gloo tests check on CPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import ops
from .preprocessing import ShiftOperator


# --------------------------------------------------------------------------------------------
# host-side plan (numpy only)
# --------------------------------------------------------------------------------------------
@dataclass
class ShardPlan:
    rank: int
    world: int
    num_nodes: int
    own: np.ndarray            # [n_own]   global ids of this rank's rows, local order
    halo: np.ndarray           # [n_halo]  global ids received, ordered by (source rank, id)
    recv_counts: np.ndarray    # [world]   halo rows coming from each rank
    send_index: np.ndarray     # [n_send]  LOCAL row indices to send, ordered by (dest rank, id)
    send_counts: np.ndarray    # [world]
    rowptr: np.ndarray         # local CSR over own rows; columns renumbered [own | halo]
    col: np.ndarray
    val: np.ndarray
    grp_rows: np.ndarray       # [n_groups_local, R] local row ids per RBU group (-1 padding)

    @property
    def n_own(self) -> int:
        return int(self.own.size)

    @property
    def n_halo(self) -> int:
        return int(self.halo.size)


def partition_rows(rowptr: np.ndarray, col: np.ndarray, val: np.ndarray, num_nodes: int,
                   world: int, R: int = 16):
    """owner[node] and the per-rank ordered row lists: contiguous ranges of the locality-ordered
    R-row groups, balanced by group count."""
    groups = ops.group_rows_host(rowptr, col, val, num_nodes, R)          # [n_groups, R]
    n_groups = groups.shape[0]
    bounds = [(n_groups * r) // world for r in range(world + 1)]
    owner = np.empty(num_nodes, np.int32)
    owned, grp_local = [], []
    for r in range(world):
        g = groups[bounds[r]:bounds[r + 1]]
        flat = g.reshape(-1)
        ids = flat[flat >= 0].astype(np.int64)
        owner[ids] = r
        owned.append(ids)
        loc = np.full(flat.shape, -1, np.int32)
        loc[flat >= 0] = np.arange(ids.size, dtype=np.int32)
        grp_local.append(loc.reshape(-1, R))
    return owner, owned, grp_local


def build_plans(rowptr: np.ndarray, col: np.ndarray, val: np.ndarray, num_nodes: int, world: int,
                R: int = 16, ranks: Optional[List[int]] = None) -> List[ShardPlan]:
    """Plans for `ranks` (default: all).  Every rank can call this with ranks=[its own rank]."""
    N = int(num_nodes)
    rowptr = np.asarray(rowptr, np.int64)
    col = np.asarray(col, np.int64)
    owner, owned, grp_local = partition_rows(rowptr, col, val, N, world, R)
    deg = np.diff(rowptr)
    row_of_e = np.repeat(np.arange(N, dtype=np.int64), deg)
    ro, co = owner[row_of_e].astype(np.int64), owner[col].astype(np.int64)
    cross = ro != co
    # distinct (receiving rank p, source rank q, node j), sorted by (p, q, j)
    key = np.unique((ro[cross] * world + co[cross]) * N + col[cross])
    p_of, q_of, j_of = key // (world * N), (key // N) % world, key % N
    plans = []
    for r in (ranks if ranks is not None else range(world)):
        own = owned[r]
        mine = p_of == r
        halo = j_of[mine]
        recv_counts = np.bincount(q_of[mine], minlength=world).astype(np.int64)
        give = q_of == r                                  # rows of mine that rank p needs
        local_of = np.full(N, -1, np.int64)
        local_of[own] = np.arange(own.size)
        send_index = local_of[j_of[give]]                 # already ordered by (p, j)
        send_counts = np.bincount(p_of[give], minlength=world).astype(np.int64)
        remap = local_of.copy()
        remap[halo] = own.size + np.arange(halo.size)
        cnt = deg[own]
        lrowptr = np.zeros(own.size + 1, np.int64)
        np.cumsum(cnt, out=lrowptr[1:])
        idx = np.repeat(rowptr[own] - lrowptr[:-1], cnt) + np.arange(int(lrowptr[-1]))
        lcol = remap[col[idx]]
        assert (lcol >= 0).all()
        plans.append(ShardPlan(r, world, N, own, halo, recv_counts, send_index.astype(np.int32),
                               send_counts, lrowptr.astype(np.int32), lcol.astype(np.int32),
                               np.asarray(val, np.float32)[idx], grp_local[r]))
    return plans


# --------------------------------------------------------------------------------------------
# device-side execution
# --------------------------------------------------------------------------------------------
class RowShardedEncoder:
    """Runs an :class:`sgp_b200.SGPEncoder` on this rank's rows of the graph."""

    def __init__(self, encoder, edge_index, edge_weight, num_nodes: int, device, group=None,
                 R: int = 16):
        self.enc, self.group, self.dev = encoder, group, torch.device(device)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        spat = encoder.sgp_encoder
        if spat.bidirectional or spat.undirected:
            raise NotImplementedError("row sharding currently covers the directed D^-1 A operator")
        from .preprocessing import build_operator
        full = build_operator(edge_index, edge_weight, num_nodes, set_diag=spat.add_self_loops,
                              device=self.dev)
        rowptr, col, val = (a.cpu().numpy() for a in full.csr_arrays())
        self.plan = build_plans(rowptr, col, val, num_nodes, self.world, R, ranks=[self.rank])[0]
        del full
        pl = self.plan
        csr = ops.Csr(torch.from_numpy(pl.rowptr).to(self.dev), torch.from_numpy(pl.col).to(self.dev),
                      torch.from_numpy(pl.val).to(self.dev), pl.n_own)
        F = encoder.reservoir.num_layers * encoder.reservoir.hidden_size
        self.op = ShiftOperator(csr, None, n_split=pl.n_own, n_cols=pl.n_own + pl.n_halo)
        if spat.rbu_mode in ("force16",) and F % 128 == 0:
            self.op.rbu = ops.rbu_build(csr, R, grp_rows_h=pl.grp_rows, n_cols=pl.n_own + pl.n_halo)
        else:
            self.op.maybe_build_rbu(F, spat.rbu_mode)
        self.send_index = torch.from_numpy(pl.send_index).to(self.dev)
        self.send_splits = [int(c) for c in pl.send_counts]
        self.recv_splits = [int(c) for c in pl.recv_counts]
        self.F = F
        self.streams = [torch.cuda.Stream(self.dev) for _ in range(2)]

    @property
    def own(self) -> np.ndarray:
        return self.plan.own

    def _exchange(self, block: torch.Tensor, send_flat: torch.Tensor, halo_flat: torch.Tensor):
        """Pack the rows other ranks need from `block` [Tc, n_own, F] and all-to-all them; returns
        the halo view [Tc, n_halo, F] (node-major storage, so every peer's segment is contiguous)."""
        Tc, _, F = block.shape
        n_send, n_halo = int(self.send_index.numel()), self.plan.n_halo
        send = send_flat[: n_send * Tc * F].view(n_send, Tc, F)
        halo = halo_flat[: n_halo * Tc * F].view(n_halo, Tc, F)
        if n_send:
            ops.gather_rows(block, self.send_index, send.permute(1, 0, 2))
        per = Tc * F
        dist.all_to_all_single(halo.view(-1), send.view(-1), [c * per for c in self.recv_splits],
                               [c * per for c in self.send_splits], group=self.group)
        return halo.permute(1, 0, 2)

    def encode_stream(self, x_own: torch.Tensor, sink: Callable[[int, int, torch.Tensor], None],
                      chunk_steps: int = 16) -> None:
        """x_own [T, n_own, Fin] (this rank's columns of the input, host or device).  `sink`
        receives each finished [t1-t0, n_own, D] device chunk on the stream it was produced on."""
        enc, pl, F, dev = self.enc, self.plan, self.F, self.dev
        res, spat = enc.reservoir, enc.sgp_encoder
        T = x_own.shape[0]
        L, H, K, D = res.num_layers, res.hidden_size, spat.receptive_field, enc.output_size
        plan = res.device_plan(dev, pl.n_own)
        state = torch.zeros(L, pl.n_own, H, device=dev)
        step = chunk_steps
        n_send = int(self.send_index.numel())
        slots = []
        for _ in range(2):
            slots.append(dict(buf=torch.empty(step, pl.n_own, D, device=dev),
                              send=torch.empty(max(n_send * step * F, 1), device=dev),
                              halo=torch.empty(max(pl.n_halo * step * F, 1), device=dev),
                              sums=torch.empty(step, F, device=dev) if spat.global_attr else None))
        main = torch.cuda.current_stream(dev)
        chunks = [(t0, min(T, t0 + step)) for t0 in range(0, T, step)]
        for s in self.streams:
            s.wait_stream(main)
        # chunks are processed in pairs: the scans are serial in time (carried state, stream 0),
        # the hop chains of the two chunks interleave on two streams.
        for i in range(0, len(chunks), 2):
            pair = chunks[i:i + 2]
            views = []
            for j, (t0, t1) in enumerate(pair):
                sl = slots[j]
                with torch.cuda.stream(self.streams[0]):
                    if j == 0:
                        self.streams[0].wait_stream(self.streams[1])   # slot 1 / state reuse
                    buf = sl["buf"][: t1 - t0]
                    xc = x_own[t0:t1].detach().to(device=dev, dtype=torch.float32, non_blocking=True)
                    res.scan_chunk(plan, xc, state, buf)
                    views.append(buf)
            self.streams[1].wait_stream(self.streams[0])
            for h in range(1, K + 1):
                for j, buf in enumerate(views):
                    with torch.cuda.stream(self.streams[j]):
                        src = buf[..., (h - 1) * F:h * F]
                        halo = self._exchange(src, slots[j]["send"], slots[j]["halo"])
                        self.op.apply(src, buf[..., h * F:(h + 1) * F], halo)
            for j, (buf, (t0, t1)) in enumerate(zip(views, pair)):
                with torch.cuda.stream(self.streams[j]):
                    if spat.global_attr:
                        sums = slots[j]["sums"][: t1 - t0]
                        ops.node_sum(buf[..., :F], sums)
                        dist.all_reduce(sums, group=self.group)
                        ops.node_mean_broadcast(sums, pl.num_nodes, buf[..., (K + 1) * F:(K + 2) * F])
                    sink(t0, t1, buf)
        for s in self.streams:
            main.wait_stream(s)


# --------------------------------------------------------------------------------------------
# bench.py, N > 1
# --------------------------------------------------------------------------------------------
def bench(args, cfg, rank, world, dev, peaks, config_dict, metric, unit, clock_sampler=None):
    """Strong scaling of the bench workload: the same N x T series, rows sharded over `world`
    GPUs.  value = N*T / max-over-ranks device time."""
    import json
    import sgp_b200
    from . import _lib
    from .synthetic import make_graph, sensor_signal

    N, T, H, K, Fin = cfg["N"], cfg["T"], cfg["H"], cfg["K"], cfg["Fin"]
    ei, ew = make_graph(cfg, seed=0)
    torch.manual_seed(2)
    enc = sgp_b200.SGPEncoder(input_size=Fin, reservoir_size=H, reservoir_layers=1, leaking_rate=0.9,
                              spectral_radius=0.9, density=0.7, input_scaling=1.0, receptive_field=K,
                              bidirectional=False, alpha_decay=False, global_attr=False)
    sh = RowShardedEncoder(enc, torch.from_numpy(ei), torch.from_numpy(ew), N, dev)
    x = sensor_signal(T, N, seed=1, exogenous=Fin == 3)
    x_host = torch.from_numpy(np.ascontiguousarray(x[:, sh.own])).pin_memory()     # this rank's rows, pinned
    x_own = x_host.to(dev)
    del x
    D = enc.output_size
    step = args.chunk or max(1, min(T, (args.chunk_mb << 20) // max(sh.plan.n_own * D * 4, 1)))
    acc = torch.zeros(1, dtype=torch.float64, device=dev)

    def one_pass():
        sh.encode_stream(x_own, lambda t0, t1, chunk: ops.checksum(chunk, acc), chunk_steps=step)

    for _ in range(args.warmup):
        one_pass()
    torch.cuda.synchronize()
    acc.zero_()
    sampler = clock_sampler(dev.index) if (clock_sampler is not None and rank == 0) else None
    if sampler is not None:
        sampler.start()
    dist.barrier()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        one_pass()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    clocks = sampler.stop() if sampler is not None else None
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # ---- end to end: every step copies this rank's rows of the series from pinned host memory
    # and reads the checksum back; the operator / halo plan is built once (it is per graph)
    chk_dev = torch.zeros(1, dtype=torch.float64, device=dev)
    chk_host = torch.zeros(1, dtype=torch.float64).pin_memory()

    def one_pass_e2e():
        xd = x_host.to(dev, non_blocking=True)
        chk_dev.zero_()
        sh.encode_stream(xd, lambda t0, t1, chunk: ops.checksum(chunk, chk_dev), chunk_steps=step)
        chk_host.copy_(chk_dev, non_blocking=True)

    one_pass_e2e()
    torch.cuda.synchronize()
    dist.barrier()
    n_e2e = max(1, min(args.steps, 2))
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(n_e2e):
        one_pass_e2e()
    f1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms_e2e = torch.tensor([f0.elapsed_time(f1) / n_e2e], device=dev)
    dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    h2d = torch.tensor([float(x_host.numel() * 4)], device=dev, dtype=torch.float64)
    dist.all_reduce(h2d)
    launches = torch.tensor([_lib.launch_count() - l0], device=dev)
    dist.all_reduce(launches)
    halo = torch.tensor([sh.plan.n_halo, sh.plan.n_own], device=dev, dtype=torch.float64)
    dist.all_reduce(halo)
    dist.all_reduce(acc)
    if rank == 0:
        ms_step = float(ms)
        value = N * T / (ms_step * 1e-3)
        line = dict(metric=metric, value=value, unit=unit, n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=ms_step, higher_is_better=True, scaling="strong",
                    vs_baseline=None, dtype="f32", data="synthetic",
                    config=config_dict(cfg, world, extra=dict(
                        chunk_steps=step, halo_rows_per_owned_row=float(halo[0] / halo[1]),
                        exchange="all_to_all_v of halo rows per hop (NCCL), 2 chunks in flight")),
                    roofline=None, cpu_baseline=None,
                    e2e=dict(value=N * T / (float(ms_e2e) * 1e-3), unit=unit, h2d_bytes_per_step=int(h2d),
                             d2h_bytes_per_step=8 * world, ms_per_step=float(ms_e2e), checksum=float(chk_host),
                             note="per step every rank copies its rows of x from pinned host memory, encodes "
                                  "(scan + halo exchange + K hops per chunk) and reads its checksum back; the "
                                  "operator and halo plan are per graph and built once, outside the step"),
                    clocks=clocks, gpu_launches=int(launches), checksum=float(acc) / args.steps)
        print(json.dumps(line))
    dist.destroy_process_group()
