"""Row-sharded encoder over the GPUs of one box (one process per GPU, torch.distributed / NCCL).

The reference is single-process; this is the B200-side scaling design named by the north star:

* the reservoir is embarrassingly parallel over nodes (shared frozen weights, per-node state) —
  every rank scans only its own rows, no communication;
* the K-hop propagation is sharded by DESTINATION rows: rank r owns one compact patch of the
  sensor graph (recursive bisection along the patch diameter, csrc/group_rows.cu::
  sgp_partition_rows — halo rows per owned row 0.05 / 0.10 / 0.19 at 2 / 4 / 8 ranks on the
  100k-node 100-NN graph) plus the matching rows of every feature block;
* before each hop the rows of the previous block that other ranks reference ("halo" rows) are
  PUSHED by one kernel (sgp_push_rows) straight into the consumers' halo buffers over NVLink:
  the buffers are torch symmetric memory mapped into every process, the kernel stores 16 bytes
  per thread to peer addresses, and a device-side barrier on either side orders it against the
  readers.  No send buffer, no second copy, no collective kernel: one pass at NVLink rate in the gap
  between two hop launches (measured at 2 GPUs: 272 ms per pass against 291 ms with pack + NCCL
  all-to-all, whose CTAs had to wait for the persistent hop CTAs to leave the SMs anyway).
  ``exchange="nccl"`` keeps the packed all-to-all-v as a fallback where peer mapping is unavailable.  The SpMM kernels read
  local columns from the rank's own block and halo columns straight from the halo buffer (second
  source pointer, no concatenation copy);
* bidirectional / undirected encoders shard the reversed / symmetrised operator by the SAME row
  partition, each with its own halo plan;
* chunks of time steps flow through a 3-stage pipeline on three CUDA streams: the scan of chunk
  c + 1 (serial in time: carried state) runs on its own stream while the hop chains of chunks c
  and c - 1 alternate on two others, so that the exchange of one overlaps the SpMM of the other;
  hand-offs are per-chunk events (scan done -> hops may start, sink done -> buffer reusable).

The partition / halo plan is numpy + one host C++ call, identical on every rank (deterministic
from the CSR), and is what the world_size-2 gloo tests check on CPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import ops
from ._lib import SgpError
from .preprocessing import ShiftOperator, build_operator, spatial_blocks


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

    @property
    def n_own(self) -> int:
        return int(self.own.size)

    @property
    def n_halo(self) -> int:
        return int(self.halo.size)


def partition_rows(rowptr: np.ndarray, col: np.ndarray, num_nodes: int, world: int) -> np.ndarray:
    """owner[node] in [0, world): compact, equally sized graph patches (host C++)."""
    if world == 1:
        return np.zeros(num_nodes, np.int32)
    return ops.partition_rows_host(rowptr, col, num_nodes, world)


def build_plans(rowptr: np.ndarray, col: np.ndarray, val: np.ndarray, num_nodes: int, world: int,
                ranks: Optional[List[int]] = None, owner: Optional[np.ndarray] = None) -> List[ShardPlan]:
    """Plans for `ranks` (default: all).  Every rank can call this with ranks=[its own rank].
    `owner` [N] fixes the row partition (a second operator over the same nodes — the reversed
    graph of a bidirectional encoder — must use the partition of the first)."""
    N = int(num_nodes)
    rowptr = np.asarray(rowptr, np.int64)
    col = np.asarray(col, np.int64)
    if owner is None:
        owner = partition_rows(rowptr, col, N, world)
    owner = np.asarray(owner, np.int64)
    deg = np.diff(rowptr)
    row_of_e = np.repeat(np.arange(N, dtype=np.int64), deg)
    ro, co = owner[row_of_e], owner[col]
    cross = ro != co
    # distinct (receiving rank p, source rank q, node j), sorted by (p, q, j)
    key = np.unique((ro[cross] * world + co[cross]) * N + col[cross])
    p_of, q_of, j_of = key // (world * N), (key // N) % world, key % N
    plans = []
    for r in (ranks if ranks is not None else range(world)):
        own = np.flatnonzero(owner == r).astype(np.int64)         # ascending global id
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
                               np.asarray(val, np.float32)[idx]))
    return plans


# --------------------------------------------------------------------------------------------
# device-side execution
# --------------------------------------------------------------------------------------------
class ShardedOperator:
    """This rank's rows of one shift operator ([own | halo] columns) + its halo exchange plan."""

    def __init__(self, plan: ShardPlan, device, F: int, rbu_mode: str = "auto"):
        self.plan = plan
        csr = ops.Csr(torch.from_numpy(plan.rowptr).to(device), torch.from_numpy(plan.col).to(device),
                      torch.from_numpy(plan.val).to(device), plan.n_own)
        self.op = ShiftOperator(csr, None, n_split=plan.n_own, n_cols=plan.n_own + plan.n_halo)
        self.op.maybe_build_rbu(F, rbu_mode)
        self.send_index = torch.from_numpy(plan.send_index).to(device)
        self.send_splits = [int(c) for c in plan.send_counts]
        self.recv_splits = [int(c) for c in plan.recv_counts]

    @property
    def n_send(self) -> int:
        return int(self.send_index.numel())


class _Phase:
    """CUDA-event pairs per phase name, on whatever stream is current (bench breakdown)."""

    def __init__(self):
        self.pairs: Dict[str, list] = {}

    def __call__(self, name):
        return _PhaseCtx(self, name)

    def totals_ms(self) -> Dict[str, float]:
        return {k: float(sum(a.elapsed_time(b) for a, b in v)) for k, v in self.pairs.items()}


class _PhaseCtx:
    def __init__(self, owner, name):
        self.owner, self.name = owner, name

    def __enter__(self):
        self.a = torch.cuda.Event(enable_timing=True)
        self.a.record()

    def __exit__(self, *exc):
        b = torch.cuda.Event(enable_timing=True)
        b.record()
        self.owner.pairs.setdefault(self.name, []).append((self.a, b))


class _NoPhase:
    def __call__(self, name):
        return self

    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return None


class PeerHalo:
    """Halo buffers in symmetric memory + per-operator destination address tables for the push."""

    def __init__(self, operators: List[ShardedOperator], F: int, step: int, device, group, n_lanes: int = 2):
        self.N_LANES = int(n_lanes)
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.F, self.step = F, step
        # identical size on every rank: the largest halo of any rank / operator
        n_max = torch.tensor([max(o.plan.n_halo for o in operators)], device=device, dtype=torch.int64)
        dist.all_reduce(n_max, op=dist.ReduceOp.MAX, group=self.group)
        self.slot = step * F                                   # floats per halo row: [step, F]
        self.lane_floats = max(int(n_max) * self.slot, 4)
        # one buffer per lane, a barrier on either side of the push.  (Two alternating buffers per lane
        # with a single barrier per hop are also correct — a push then only overwrites rows read two
        # hops ago — but measured 3-4 % slower at 2 and 8 GPUs: the "buffer free" barrier also paces the
        # pushes into the gaps between hop launches.)
        self.n_bufs = self.N_LANES
        self.buf = symm_mem.empty(self.n_bufs * self.lane_floats, dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        # every rank's receive layout: recv_counts[p][q] = rows rank p receives from q (its halo is
        # ordered by source rank), so my rows for p start at slot sum(recv_counts[p][:me])
        self.addr = []                                          # [operator][lane] -> int64 [n_send] (device)
        for o in operators:
            mine = torch.tensor(o.plan.recv_counts, device=device, dtype=torch.int64)
            allc = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allc, mine, group=self.group)
            allc = torch.stack(allc).cpu().numpy()              # [p, q]
            per_lane = []
            for lane in range(self.n_bufs):
                parts = []
                for p in range(world):
                    n = int(o.plan.send_counts[p])
                    first = int(allc[p][:rank].sum())
                    assert n == int(allc[p][rank])
                    base = ptrs[p] + 4 * lane * self.lane_floats
                    parts.append(base + 4 * self.slot * (first + np.arange(n, dtype=np.int64)))
                a = np.concatenate(parts) if parts else np.zeros(0, np.int64)
                per_lane.append(torch.from_numpy(a).to(device))
            self.addr.append(per_lane)

    def halo_view(self, b: int, n_halo: int, Tc: int) -> torch.Tensor:
        """[Tc, n_halo, F] view of this rank's buffer b (node-major slots of `step` x F floats)."""
        flat = self.buf[b * self.lane_floats: b * self.lane_floats + max(n_halo, 1) * self.slot]
        return flat.view(max(n_halo, 1), self.step, self.F)[:n_halo, :Tc].permute(1, 0, 2)

    def barrier(self, lane: int) -> None:
        # bounded: a peer that never arrives traps the kernel (CUDA error) instead of hanging the box
        self.hdl.barrier(channel=lane, timeout_ms=20000)


class RowShardedEncoder:
    """Runs an :class:`sgp_b200.SGPEncoder` on this rank's rows of the graph."""

    def __init__(self, encoder, edge_index, edge_weight, num_nodes: int, device, group=None,
                 exchange: str = "auto", lanes: int = 2):
        """exchange: "p2p" (halo rows pushed into peer-mapped buffers by sgp_push_rows), "nccl"
        (pack + all_to_all_single), "auto" = p2p when the symmetric-memory rendezvous works.
        lanes: hop chains in flight (streams); lanes + 1 chunk buffers."""
        self.enc, self.group, self.dev = encoder, group, torch.device(device)
        self.N_LANES, self.N_SLOTS = int(lanes), int(lanes) + 1
        self.exchange_mode, self._peer, self._peer_key = exchange, None, None
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        # SMs left free for the halo push / barrier kernels (they cannot share an SM with a hop CTA)
        import os
        # measured on C4 (profiles/r2_bench_n8_p2p_free*.json, n4): 12 free SMs give 1447 M against 1314 M
        # node-steps/s at 8 GPUs and 723 M against 707 M at 4; at 2 GPUs the hop is too long for it to pay
        free = os.environ.get("SGP_B200_FREE_SMS")
        self.free_sms = int(free) if free is not None else (12 if self.world >= 4 else 0)
        ops.tc_set_cta_limit(148 - self.free_sms)
        spat = encoder.sgp_encoder
        if spat.undirected:
            assert spat.bidirectional is False
        F = encoder.reservoir.num_layers * encoder.reservoir.hidden_size
        self.F = F
        # the full operators, exactly as SGPSpatialEncoder builds them (make_operators), then cut
        full = build_operator(edge_index, edge_weight, num_nodes, gcn_norm=spat.undirected,
                              set_diag=spat.add_self_loops, symmetrize=spat.undirected, device=self.dev)
        rowptr, col, val = ops.to_host(*full.csr_arrays())
        del full
        self.owner = partition_rows(rowptr, col, num_nodes, self.world)
        self.plan = build_plans(rowptr, col, val, num_nodes, self.world, ranks=[self.rank],
                                owner=self.owner)[0]
        self.fwd = ShardedOperator(self.plan, self.dev, F, spat.rbu_mode)
        self.bwd: Optional[ShardedOperator] = None
        if spat.bidirectional:
            rev = build_operator(edge_index, edge_weight, num_nodes, gcn_norm=False,
                                 set_diag=spat.add_self_loops, transpose=True, device=self.dev)
            rowptr, col, val = ops.to_host(*rev.csr_arrays())
            del rev
            plan_b = build_plans(rowptr, col, val, num_nodes, self.world, ranks=[self.rank],
                                 owner=self.owner)[0]
            assert np.array_equal(plan_b.own, self.plan.own)
            self.bwd = ShardedOperator(plan_b, self.dev, F, spat.rbu_mode)
        self.op = self.fwd.op                                   # kept: the forward operator
        self.s_scan = torch.cuda.Stream(self.dev)
        self.s_hop = [torch.cuda.Stream(self.dev) for _ in range(self.N_LANES)]

    @property
    def own(self) -> np.ndarray:
        return self.plan.own

    @property
    def operators(self) -> List[ShardedOperator]:
        return [o for o in (self.fwd, self.bwd) if o is not None]

    def halo_rows(self) -> int:
        return sum(o.plan.n_halo for o in self.operators)

    def _peer_halo(self, step: int) -> Optional[PeerHalo]:
        """The symmetric halo buffers for chunks of `step` time steps (built once per step size;
        collective).  None = use the NCCL exchange."""
        if self.exchange_mode == "nccl" or self.world == 1 or self.F % 4:
            return None
        if self._peer_key != step:
            ok = 1
            try:
                self._peer = PeerHalo(self.operators, self.F, step, self.dev, self.group, self.N_LANES)
            except Exception as e:  # noqa: BLE001 - no peer mapping on this box: fall back together
                if self.exchange_mode == "p2p":
                    raise
                self._peer, self._peer_error, ok = None, repr(e), 0
            flag = torch.tensor([ok], device=self.dev, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            if int(flag.item()) == 0:
                self._peer = None
            self._peer_key = step
        return self._peer

    def _exchange_p2p(self, peer: PeerHalo, oi: int, sop: ShardedOperator, block: torch.Tensor, lane: int, phase):
        """barrier (every rank is done reading this lane's halo buffer) -> push my rows into the
        consumers' buffers -> barrier (every push has landed); returns my halo view."""
        Tc = block.shape[0]
        with phase("exchange"):
            peer.barrier(lane)
        with phase("pack"):
            if sop.n_send:
                ops.push_rows(block, sop.send_index, peer.addr[oi][lane], self.F)
        with phase("exchange"):
            peer.barrier(lane)
        return peer.halo_view(lane, sop.plan.n_halo, Tc)

    def _exchange(self, sop: ShardedOperator, block: torch.Tensor, send_flat: torch.Tensor,
                  halo_flat: torch.Tensor, phase):
        """Pack the rows other ranks need from `block` [Tc, n_own, F] and all-to-all them; returns
        the halo view [Tc, n_halo, F] (node-major storage, so every peer's segment is contiguous)."""
        Tc, _, F = block.shape
        n_send, n_halo = sop.n_send, sop.plan.n_halo
        send = send_flat[: n_send * Tc * F].view(n_send, Tc, F)
        halo = halo_flat[: n_halo * Tc * F].view(n_halo, Tc, F)
        with phase("pack"):
            if n_send:
                ops.gather_rows(block, sop.send_index, send.permute(1, 0, 2))
        per = Tc * F
        with phase("exchange"):
            dist.all_to_all_single(halo.view(-1), send.view(-1), [c * per for c in sop.recv_splits],
                                   [c * per for c in sop.send_splits], group=self.group)
        return halo.permute(1, 0, 2)

    def encode_stream(self, x_own: torch.Tensor, sink: Optional[Callable[[int, int, torch.Tensor], None]],
                      chunk_steps: int = 16, checksum: Optional[torch.Tensor] = None,
                      breakdown: Optional[dict] = None) -> None:
        """x_own [T, n_own, Fin] (this rank's columns of the input, host or device).  `sink`
        receives each finished [t1-t0, n_own, D] device chunk on the stream it was produced on
        (may be None when only `checksum` — a device float64 scalar that receives the sum of the
        rank's whole output, accumulated by the producing kernels — is wanted).  `breakdown`, when
        given, is filled with per-phase stream-busy milliseconds (CUDA events; phases on different
        streams overlap, so they add up to more than the wall time)."""
        enc, pl, F, dev = self.enc, self.plan, self.F, self.dev
        res, spat = enc.reservoir, enc.sgp_encoder
        T = x_own.shape[0]
        L, H, K, D = res.num_layers, res.hidden_size, spat.receptive_field, enc.output_size
        plan = res.device_plan(dev, pl.n_own)
        state = torch.zeros(L, pl.n_own, H, device=dev)
        step = max(1, int(chunk_steps))
        phase = _Phase() if breakdown is not None else _NoPhase()
        n_send = max(o.n_send for o in self.operators)
        n_halo = max(o.plan.n_halo for o in self.operators)
        bufs = [torch.empty(step, pl.n_own, D, device=dev) for _ in range(self.N_SLOTS)]
        peer = self._peer_halo(step)
        self.exchange_used = "p2p push (symmetric memory)" if peer is not None else "nccl all_to_all_v"
        lanes = [dict(send=None if peer is not None else torch.empty(max(n_send * step * F, 1), device=dev),
                      halo=None if peer is not None else torch.empty(max(n_halo * step * F, 1), device=dev),
                      sums=torch.empty(step, F, device=dev) if spat.global_attr else None)
                 for _ in range(self.N_LANES)]
        main = torch.cuda.current_stream(dev)
        chunks = [(t0, min(T, t0 + step)) for t0 in range(0, T, step)]
        for s in (self.s_scan, *self.s_hop):
            s.wait_stream(main)
        slot_free: List[Optional[torch.cuda.Event]] = [None] * self.N_SLOTS
        g = spatial_blocks(K, spat.bidirectional)
        for c, (t0, t1) in enumerate(chunks):
            slot, lane = c % self.N_SLOTS, c % self.N_LANES
            buf = bufs[slot][: t1 - t0]
            # ---- scan: serial in time (carried state), its own stream ----
            with torch.cuda.stream(self.s_scan):
                if slot_free[slot] is not None:
                    self.s_scan.wait_event(slot_free[slot])
                xc = x_own[t0:t1].detach().to(device=dev, dtype=torch.float32, non_blocking=True)
                with phase("scan"):
                    res.scan_chunk(plan, xc, state, buf, checksum)
                scan_done = torch.cuda.Event()
                scan_done.record()
            # ---- hop chains: alternate between two streams ----
            hs = self.s_hop[lane]
            with torch.cuda.stream(hs):
                hs.wait_event(scan_done)
                bounds = [res.state_bound(), res.state_bound()]
                for oi, (sop, base) in enumerate(((self.fwd, 0), (self.bwd, K))):
                    if sop is None:
                        continue
                    for h in range(1, K + 1):
                        src = buf[..., :F] if h == 1 else buf[..., (base + h - 1) * F:(base + h) * F]
                        if peer is not None:
                            halo = self._exchange_p2p(peer, oi, sop, src, lane, phase)
                        else:
                            halo = self._exchange(sop, src, lanes[lane]["send"], lanes[lane]["halo"], phase)
                        with phase("hop"):
                            sop.op.apply(src, buf[..., (base + h) * F:(base + h + 1) * F], halo, checksum,
                                         bound=bounds[oi])
                            bounds[oi] = sop.op.out_bound(bounds[oi])
                if spat.global_attr:
                    with phase("global"):
                        sums = lanes[lane]["sums"][: t1 - t0]
                        ops.node_sum(buf[..., :F], sums)
                        dist.all_reduce(sums, group=self.group)
                        ops.node_mean_broadcast(sums, pl.num_nodes, buf[..., g * F:(g + 1) * F])
                        if checksum is not None:
                            ops.checksum_view(buf[..., g * F:(g + 1) * F], checksum)
                if sink is not None:
                    with phase("sink"):
                        sink(t0, t1, buf)
                ev = torch.cuda.Event()
                ev.record()
                slot_free[slot] = ev
        for s in (self.s_scan, *self.s_hop):
            main.wait_stream(s)
        if breakdown is not None:
            torch.cuda.synchronize(dev)
            breakdown.update(phase.totals_ms())
        self._plan_for_check = plan

    def check(self) -> None:
        """Raise on EVERY rank if a tensor-core launch of any rank reported a barrier timeout
        (results invalid).  Synchronises the device and the group."""
        bad = 0
        for sop in self.operators:
            for fmt in (sop.op.tc, sop.op.tc16):
                if fmt is not None:
                    bad |= int(fmt.err.item() != 0)
        for entry in getattr(self, "_plan_for_check", []) or []:
            if entry[0] in ("tc", "tc16"):
                bad |= int(entry[-1].item() != 0)
        flag = torch.tensor([bad], device=self.dev, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
        if int(flag.item()) != 0:
            raise SgpError("row-sharded encode: a tensor-core kernel reported a barrier timeout on at "
                           "least one rank (results invalid)")


def encode_sharded_lockstep(encoder, edge_index, edge_weight, num_nodes: int, x: torch.Tensor,
                            world: int, device) -> torch.Tensor:
    """Single-process, single-GPU emulation of the row-sharded encode: `world` shards are built
    with the same plans and device kernels (halo pack, second-source SpMM) and advanced hop by hop
    in lockstep, the all-to-all replaced by direct copies between the shards' buffers.  Returns the
    assembled [T, N, D] output.  This is what the one-GPU parity test drives; the NCCL pipeline
    itself is checked by tools/check_sharded.py under torchrun."""
    dev = torch.device(device)
    spat, res = encoder.sgp_encoder, encoder.reservoir
    F, K, D = res.num_layers * res.hidden_size, spat.receptive_field, encoder.output_size
    T = x.shape[0]
    specs = [dict(gcn_norm=spat.undirected, set_diag=spat.add_self_loops, symmetrize=spat.undirected)]
    if spat.bidirectional:
        specs.append(dict(gcn_norm=False, set_diag=spat.add_self_loops, transpose=True))
    owner, shards = None, []           # shards[o][r] = ShardedOperator
    for spec in specs:
        full = build_operator(edge_index, edge_weight, num_nodes, device=dev, **spec)
        rowptr, col, val = ops.to_host(*full.csr_arrays())
        if owner is None:
            owner = partition_rows(rowptr, col, num_nodes, world)
        plans = build_plans(rowptr, col, val, num_nodes, world, owner=owner)
        shards.append([ShardedOperator(p, dev, F, spat.rbu_mode) for p in plans])
    plans = [s.plan for s in shards[0]]
    xs = x.detach().to(device=dev, dtype=torch.float32)
    bufs = []
    for p in plans:
        buf = torch.empty(T, p.n_own, D, device=dev)
        state = torch.zeros(res.num_layers, p.n_own, res.hidden_size, device=dev)
        plan = res.device_plan(dev, p.n_own)
        res.scan_chunk(plan, xs[:, torch.from_numpy(p.own).to(dev)].contiguous(), state, buf)
        res.check_plan(plan)
        bufs.append(buf)
    for o, base in zip(range(len(shards)), (0, K)):
        bound = res.state_bound()
        for h in range(1, K + 1):
            sl = slice(0, F) if h == 1 else slice((base + h - 1) * F, (base + h) * F)
            # pack on every shard, then "exchange": rank p's halo segment from q = q's send segment for p
            sends = []
            for r, sop in enumerate(shards[o]):
                send = torch.empty(sop.n_send, T, F, device=dev)
                if sop.n_send:
                    ops.gather_rows(bufs[r][..., sl], sop.send_index, send.permute(1, 0, 2))
                sends.append(send)
            for r, sop in enumerate(shards[o]):
                parts = []
                for q, sq in enumerate(shards[o]):
                    off = int(np.sum(sq.plan.send_counts[:r]))
                    parts.append(sends[q][off:off + int(sq.plan.send_counts[r])])
                halo = torch.cat(parts, 0) if parts else torch.empty(0, T, F, device=dev)
                assert halo.shape[0] == sop.plan.n_halo
                sop.op.apply(bufs[r][..., sl], bufs[r][..., (base + h) * F:(base + h + 1) * F],
                             halo.permute(1, 0, 2), bound=bound)
                sop.op.check()
            bound = max((s_.op.out_bound(bound) or 0.0) for s_ in shards[o]) or None
    out = torch.empty(T, num_nodes, D, device=dev)
    if spat.global_attr:
        g = spatial_blocks(K, spat.bidirectional)
        total = torch.zeros(T, F, device=dev)
        for r in range(world):
            sums = torch.empty(T, F, device=dev)
            ops.node_sum(bufs[r][..., :F], sums)
            total += sums
        for r in range(world):
            ops.node_mean_broadcast(total, num_nodes, bufs[r][..., g * F:(g + 1) * F])
    for r, p in enumerate(plans):
        out[:, torch.from_numpy(p.own).to(dev)] = bufs[r]
    return out


# --------------------------------------------------------------------------------------------
# bench.py, N > 1
# --------------------------------------------------------------------------------------------
def bench(args, cfg, rank, world, dev, peaks, config_dict, metric, unit, clock_sampler=None):
    """Strong scaling of the bench workload: the same N x T series, rows sharded over `world`
    GPUs.  value = N*T / max-over-ranks device time."""
    import json
    import time
    import sgp_b200
    from . import _lib
    from .synthetic import make_graph, sensor_signal

    N, T, H, K, Fin = cfg["N"], cfg["T"], cfg["H"], cfg["K"], cfg["Fin"]
    ei, ew = make_graph(cfg, seed=0)
    ei_t, ew_t = torch.from_numpy(ei), torch.from_numpy(ew)
    torch.manual_seed(2)
    enc = sgp_b200.SGPEncoder(input_size=Fin, reservoir_size=H, reservoir_layers=1, leaking_rate=0.9,
                              spectral_radius=0.9, density=0.7, input_scaling=1.0, receptive_field=K,
                              bidirectional=False, alpha_decay=False, global_attr=False)
    torch.cuda.synchronize()
    t_b0 = time.perf_counter()
    import os
    sh = RowShardedEncoder(enc, ei_t, ew_t, N, dev, exchange=os.environ.get("SGP_B200_EXCHANGE", "auto"),
                           lanes=int(os.environ.get("SGP_B200_LANES", 2)))
    torch.cuda.synchronize()
    build_ms = torch.tensor([(time.perf_counter() - t_b0) * 1e3], device=dev)
    dist.all_reduce(build_ms, op=dist.ReduceOp.MAX)
    x = sensor_signal(T, N, seed=1, exogenous=Fin == 3)
    x_host = torch.from_numpy(np.ascontiguousarray(x[:, sh.own])).pin_memory()     # this rank's rows, pinned
    x_own = x_host.to(dev)
    del x
    D = enc.output_size
    # >= 32 chunks per pass so that the 3-stage pipeline has something to overlap
    from .preprocessing import round_chunk_steps
    step = args.chunk or round_chunk_steps(min((T + 31) // 32, (args.chunk_mb << 20) // max(sh.plan.n_own * D * 4, 1)), T)
    acc = torch.zeros(1, dtype=torch.float64, device=dev)

    def one_pass(breakdown=None):
        sh.encode_stream(x_own, None, chunk_steps=step, checksum=acc, breakdown=breakdown)

    sampler = clock_sampler(dev.index) if (clock_sampler is not None and rank == 0) else None
    if sampler is not None:
        sampler.start()                  # from the warm-up on: a timed pass is tens of ms, the sampler ticks every 50 ms
    for _ in range(args.warmup):
        one_pass()
    torch.cuda.synchronize()
    sh.check()
    acc.zero_()
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
    launches = torch.tensor([_lib.launch_count() - l0], device=dev)
    sh.check()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    chk_timed = acc.clone()
    dist.all_reduce(chk_timed)
    # ---- one more, instrumented pass: per-phase stream-busy time on this rank (max over ranks) ----
    bd: dict = {}
    one_pass(bd)
    names = ["scan", "pack", "exchange", "hop", "global", "sink"]
    bd_t = torch.tensor([bd.get(k, 0.0) for k in names], device=dev, dtype=torch.float64)
    bd_max, bd_sum = bd_t.clone(), bd_t.clone()
    dist.all_reduce(bd_max, op=dist.ReduceOp.MAX)
    dist.all_reduce(bd_sum)
    # ---- end to end: every step copies this rank's rows of the series from pinned host memory
    # and reads the checksum back; the operator / halo plan is per graph and built once (build_ms)
    chk_dev = torch.zeros(1, dtype=torch.float64, device=dev)
    chk_host = torch.zeros(1, dtype=torch.float64).pin_memory()

    def one_pass_e2e():
        xd = x_host.to(dev, non_blocking=True)
        chk_dev.zero_()
        sh.encode_stream(xd, None, chunk_steps=step, checksum=chk_dev)
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
    sh.check()
    ms_e2e = torch.tensor([f0.elapsed_time(f1) / n_e2e], device=dev)
    dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    h2d = torch.tensor([float(x_host.numel() * 4)], device=dev, dtype=torch.float64)
    dist.all_reduce(h2d)
    dist.all_reduce(launches)
    halo = torch.tensor([sh.halo_rows(), sh.plan.n_own], device=dev, dtype=torch.float64)
    dist.all_reduce(halo)
    chk_e2e = torch.tensor([float(chk_host)], device=dev, dtype=torch.float64)
    dist.all_reduce(chk_e2e)
    # roofline of the hop on the shards: algorithmic bytes of every rank's launches (its rows of the
    # operator once per launch + its rows of the panel read and written once per time step; halo rows
    # are traffic, not algorithmic bytes) over the slowest rank's hop time in the instrumented pass
    n_chunks = (T + step - 1) // step
    local_bytes = torch.tensor([float(n_chunks * K * (8 * sh.fwd.op.csr.nnz + 4 * (sh.plan.n_own + 1)) +
                                      K * 2 * T * sh.plan.n_own * H * 4)], device=dev, dtype=torch.float64)
    dist.all_reduce(local_bytes)
    if rank == 0:
        ms_step = float(ms)
        value = N * T / (ms_step * 1e-3)
        fmt = sh.fwd.op
        hop_ms = float(bd_max[names.index("hop")])
        achieved = float(local_bytes) / (hop_ms * 1e-3) / 1e9 if hop_ms else 0.0
        roofline = dict(bound="hbm", kernel=("spmm_rbu_tc16_kernel<HALO> on %d shards" % world if fmt.tc16 is not None else
                                             "spmm_rbu_tc_kernel<HALO> on %d shards" % world if fmt.tc is not None else "spmm (CUDA cores)"),
                        achieved=achieved, peak=peaks["hbm_gbs"] * world, unit="GB/s",
                        frac=achieved / (peaks["hbm_gbs"] * world), peak_source=peaks["source"] + " x n_gpus",
                        traffic=None, algorithmic_bytes_per_pass_all_ranks=float(local_bytes),
                        hop_ms_per_pass_slowest_rank=hop_ms,
                        note="aggregate over ranks: all ranks' algorithmic hop bytes / the slowest rank's hop "
                             "stream time (instrumented pass, CUDA events)")
        nvlink_bytes = float(halo[0]) * 4 * H * K * T          # halo rows x row bytes x hops x time steps
        line = dict(metric=metric, value=value, unit=unit, n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=ms_step, higher_is_better=True, scaling="strong",
                    vs_baseline=None, dtype="f32", data="synthetic",
                    config=config_dict(cfg, world),
                    kernel_config=dict(
                        chunk_steps=step, halo_rows_per_owned_row=float(halo[0] / halo[1]),
                        operator_format=("tcgen05 fp16x3 96-row groups" if fmt.tc16 is not None else "tcgen05 64-row groups" if fmt.tc is not None else
                                         "rbu%d" % fmt.rbu.R if fmt.rbu is not None else "csr"),
                        partition="recursive bisection along the patch diameter (sgp_partition_rows)",
                        exchange=sh.exchange_used + " of halo rows per hop; scan + %d hop chains on %d streams, "
                                 "%d chunk buffers; %d SMs kept free of hop CTAs" % (sh.N_LANES, sh.N_LANES + 1, sh.N_SLOTS, sh.free_sms),
                        sink="fp64 checksum of the whole output, accumulated in the scan / hop epilogues"),
                    roofline=roofline,
                    breakdown=dict(
                        unit="ms of stream-busy time per pass, CUDA events on the phase's own stream; phases on "
                             "different streams overlap (sum > ms_per_step); 'exchange' includes waiting for peers",
                        max_over_ranks={k: float(v) for k, v in zip(names, bd_max)},
                        mean_over_ranks={k: float(v) / world for k, v in zip(names, bd_sum)},
                        nvlink_bytes_received_per_pass_all_ranks=nvlink_bytes,
                        operator_build_ms_max_over_ranks=float(build_ms)),
                    cpu_baseline=None,
                    e2e=dict(value=N * T / (float(ms_e2e) * 1e-3), unit=unit, h2d_bytes_per_step=int(h2d),
                             d2h_bytes_per_step=8 * world, ms_per_step=float(ms_e2e), checksum=float(chk_e2e),
                             note="per step every rank copies its rows of x from pinned host memory, encodes "
                                  "(scan + halo exchange + K hops per chunk, checksum fused into the kernels' "
                                  "epilogues) and reads its checksum back; the operator and halo plan are per "
                                  "graph, built once outside the step (operator_build_ms) — as in the 1-GPU line"),
                    clocks=clocks, gpu_launches=int(launches), checksum=float(chk_timed) / args.steps)
        print(json.dumps(line))
    dist.destroy_process_group()
