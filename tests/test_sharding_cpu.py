"""Host logic of the row-sharded path on CPU: partition (host C++) / halo plan (numpy) and the
all-to-all-v exchange pattern with world_size 2 and 3 over gloo.  The device kernels are not
involved; the SpMM check uses scipy on the plan's local CSR."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sgp_oracle as O
from sgp_b200.sharded import build_plans, partition_rows
from sgp_b200.synthetic import sensor_knn
from tests.helpers import random_graph


def _global_csr(n, kind):
    if kind == "knn":
        ei, ew = sensor_knn(n, 12, seed=4)
    else:
        ei, ew = random_graph(n, 9 * n, seed=4)
    return O.build_operator(ei, ew, n, set_diag=False)


@pytest.mark.parametrize("kind", ["knn", "random"])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_plans_cover_rows_and_reproduce_spmm(kind, world):
    n, F = 803, 5
    rowptr, col, val = _global_csr(n, kind)
    plans = build_plans(rowptr, col, val, n, world)
    assert sorted(np.concatenate([p.own for p in plans]).tolist()) == list(range(n))
    X = np.random.default_rng(0).standard_normal((n, F)).astype(np.float32)
    S = sp.csr_matrix((val, col, rowptr), shape=(n, n))
    want = S @ X
    for p in plans:
        assert not np.intersect1d(p.own, p.halo).size
        assert p.recv_counts.sum() == p.halo.size and p.send_counts.sum() == p.send_index.size
        assert p.recv_counts[p.rank] == 0 and p.send_counts[p.rank] == 0
        local = sp.csr_matrix((p.val, p.col, p.rowptr), shape=(p.n_own, p.n_own + p.n_halo))
        src = np.concatenate([X[p.own], X[p.halo]])
        np.testing.assert_allclose(local @ src, want[p.own], rtol=1e-5, atol=1e-6)
        assert abs(p.n_own - n / world) <= 1                          # balanced patches
    # what rank q sends to p is exactly p's halo segment for q, in the same order
    for p in plans:
        off = np.concatenate([[0], np.cumsum(p.recv_counts)])
        for q in plans:
            soff = np.concatenate([[0], np.cumsum(q.send_counts)])
            sent = q.own[q.send_index[soff[p.rank]:soff[p.rank + 1]]]
            np.testing.assert_array_equal(sent, p.halo[off[q.rank]:off[q.rank + 1]])
    if kind == "knn" and world == 2:
        assert sum(p.n_halo for p in plans) < 0.5 * n       # locality-aware: small halo


def test_second_operator_uses_the_first_partition():
    """Bidirectional encoders shard the reversed operator by the forward operator's row partition."""
    n = 640
    ei, ew = sensor_knn(n, 10, seed=2)
    fwd = O.build_operator(ei, ew, n, set_diag=False)
    bwd = O.build_operator(ei[[1, 0]], ew, n, set_diag=False)
    owner = partition_rows(fwd[0], fwd[1], n, 4)
    pf = build_plans(*fwd, n, 4, owner=owner)
    pb = build_plans(*bwd, n, 4, owner=owner)
    X = np.random.default_rng(3).standard_normal((n, 3)).astype(np.float32)
    Sb = sp.csr_matrix((bwd[2], bwd[1], bwd[0]), shape=(n, n))
    for a, b in zip(pf, pb):
        np.testing.assert_array_equal(a.own, b.own)
        local = sp.csr_matrix((b.val, b.col, b.rowptr), shape=(b.n_own, b.n_own + b.n_halo))
        np.testing.assert_allclose(local @ np.concatenate([X[b.own], X[b.halo]]), (Sb @ X)[b.own],
                                   rtol=1e-5, atol=1e-6)


def test_partition_halo_fraction_at_bench_size():
    """BASELINE C4 graph (N = 100k, 100-NN): compact patches keep the halo at <= 0.20 rows per owned
    row on 8 ranks (contiguous ranges of one breadth-first order gave 0.318 in round 1)."""
    n, k = 100_000, 100
    ei, ew = sensor_knn(n, k, seed=0)
    rowptr, col, _ = O.build_operator(ei, ew, n, set_diag=False)
    deg = np.diff(rowptr)
    row = np.repeat(np.arange(n), deg)
    want = {2: 0.06, 4: 0.13, 8: 0.20}
    for world, limit in want.items():
        owner = partition_rows(rowptr, col, n, world)
        assert np.bincount(owner, minlength=world).tolist() == [n // world] * world
        cross = owner[row] != owner[col]
        halo = np.unique(owner[row][cross].astype(np.int64) * n + col[cross]).size
        assert halo / n <= limit, (world, halo / n)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, F, Tc, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rowptr, col, val = _global_csr(n, "knn")
        plan = build_plans(rowptr, col, val, n, world, ranks=[rank])[0]
        X = torch.from_numpy(np.random.default_rng(1).standard_normal((Tc, n, F)).astype(np.float32))
        own = torch.from_numpy(plan.own)
        block = X[:, own]                                          # [Tc, n_own, F]
        # node-major packing, exactly like RowShardedEncoder._exchange
        send = block[:, torch.from_numpy(plan.send_index).long()].permute(1, 0, 2).contiguous()
        halo = torch.empty(plan.n_halo, Tc, F)
        per = Tc * F
        dist.all_to_all_single(halo.view(-1), send.view(-1), [int(c) * per for c in plan.recv_counts],
                               [int(c) * per for c in plan.send_counts])
        ok = torch.equal(halo.permute(1, 0, 2), X[:, torch.from_numpy(plan.halo)])
        S = sp.csr_matrix((val, col, rowptr), shape=(n, n))
        local = sp.csr_matrix((plan.val, plan.col, plan.rowptr), shape=(plan.n_own, plan.n_own + plan.n_halo))
        for t in range(Tc):
            src = np.concatenate([block[t].numpy(), halo[:, t].numpy()])
            ok = ok and np.allclose(local @ src, (S @ X[t].numpy())[plan.own], rtol=1e-5, atol=1e-6)
        flag = torch.tensor([1.0 if ok else 0.0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret.put(float(flag))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_over_gloo(world):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 400, 6, 3, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=10) == 1.0


def test_row_sharded_encoder_constructor_on_cpu(monkeypatch, tmp_path):
    """RowShardedEncoder.__init__ end to end on the CPU (world_size 1 over gloo), with the device pieces
    stubbed: partition, per-operator plans of a bidirectional encoder, free-SM default, stream set-up."""
    import sgp_b200
    from sgp_b200 import sharded

    class _FullOp:
        def __init__(self, arrs):
            self.arrs = arrs

        def csr_arrays(self):
            return tuple(torch.from_numpy(np.ascontiguousarray(a)) for a in self.arrs)

    def fake_build_operator(edge_index, edge_weight, num_nodes, *, gcn_norm=False, set_diag=False, remove_diag=False,
                            symmetrize=False, transpose=False, normalize=True, device=None):
        ei = np.asarray(edge_index)
        if transpose:
            ei = ei[[1, 0]]
        return _FullOp(O.build_operator(ei, np.asarray(edge_weight), num_nodes, gcn_norm=gcn_norm, set_diag=set_diag))

    class _StubShardedOperator:
        def __init__(self, plan, device, F, rbu_mode="auto"):
            self.plan, self.op = plan, None

    limits = []
    monkeypatch.setattr(sharded, "build_operator", fake_build_operator)
    monkeypatch.setattr(sharded, "ShardedOperator", _StubShardedOperator)
    monkeypatch.setattr(sharded.ops, "tc_set_cta_limit", limits.append)
    monkeypatch.setattr(torch.cuda, "Stream", lambda *a, **k: object())
    dist.init_process_group("gloo", init_method=f"file://{tmp_path}/rdzv", rank=0, world_size=1)
    try:
        n = 300
        ei, ew = sensor_knn(n, 8, seed=1)
        torch.manual_seed(0)
        enc = sgp_b200.SGPEncoder(3, 16, 1, 0.9, 0.9, 0.7, 1.0, 2, True, False, True)
        sh = sharded.RowShardedEncoder(enc, torch.from_numpy(ei), torch.from_numpy(ew), n, "cpu")
        assert sh.world == 1 and sh.free_sms == 0 and limits == [148]
        assert sh.bwd is not None and np.array_equal(sh.bwd.plan.own, sh.fwd.plan.own) and sh.plan.n_own == n
        assert sh.halo_rows() == 0 and len(sh.operators) == 2 and len(sh.s_hop) == 2
    finally:
        dist.destroy_process_group()
