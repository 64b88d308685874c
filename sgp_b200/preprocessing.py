"""Spatial propagation — host side of kernels K2/K3/K4.

Same function names and arguments as the reference's ``lib/sgp_preprocessing.py``
(``preprocess_adj``, ``sgp_spatial_embedding``, ``reservoir_preprocessing_``,
``preprocess_dataset``); the sparse algebra that the reference delegates to torch_sparse /
torch_geometric runs in the sm_100a kernels behind ``include/sgp_b200.h``:

* ``preprocess_adj`` returns a :class:`ShiftOperator` (device CSR, optionally also the RBU
  format) where the reference returns a ``torch_sparse.SparseTensor``; it supports ``op @ x``.
* ``sgp_spatial_embedding`` fills one concatenated ``[B, N, (1+K')F]`` buffer hop by hop (no
  ``torch.cat``) and returns the reference's list as views of it.
"""
from __future__ import annotations

import math
import os
from typing import List, Optional, Union

import numpy as np
import torch
from torch import Tensor

from . import _lib, ops
from ._lib import SgpError
from .reservoir import Reservoir, _cuda_device_for

# rows are grouped for the RBU kernel when the graph is big enough for gather traffic to matter
_RBU_MIN_NNZ = int(os.environ.get("SGP_B200_RBU_MIN_NNZ", 200_000))
_RBU_MIN_FILL = {16: 0.30, 8: 0.40, 4: 0.55}
_TC_MIN_FILL = 0.10      # 64-row groups: the tensor-core hop wins as long as the slabs are >= 10 % dense
# tensor-core hop format chosen by "auto": "tc16" (fp16x3, 96-row groups: 75.5 us per hop-panel at C4) or
# "tc" (3xTF32, 64-row groups: 92.1 us); tc16 needs a bound on |x| per launch — the encoders know it (tanh
# states), the generic entry points take max|x| of their input
_HOP_DEFAULT = os.environ.get("SGP_B200_HOP", "tc16")


class ShiftOperator:
    """Normalised graph-shift operator resident on one GPU (what ``preprocess_adj`` returns)."""

    def __init__(self, csr: ops.Csr, rbu: Optional[ops.Rbu] = None, n_split: int = 0,
                 tc: Optional[ops.TcOp] = None, n_cols: Optional[int] = None,
                 tc16: Optional[ops.Tc16Op] = None):
        self.csr = csr
        self.rbu = rbu
        self.tc = tc
        self.tc16 = tc16            # fp16x3 / 96-row format: used when the caller knows a bound on |x|
        self.n_cols = n_cols        # > num rows for a row-sharded (local + halo columns) operator
        self.num_nodes = csr.num_nodes
        self.n_split = n_split      # row-sharded: column ids >= n_split address the halo buffer

    @property
    def device(self):
        return self.csr.rowptr.device

    def sparse_sizes(self):
        return (self.num_nodes, self.num_nodes)

    def nnz(self) -> int:
        return self.csr.nnz

    def csr_arrays(self):
        """(rowptr, col, val) as they would come out of SparseTensor.csr() (int32 indices)."""
        return self.csr.rowptr, self.csr.col, self.csr.val

    def maybe_build_rbu(self, F: int, mode: str = "auto") -> None:
        """Attach a grouped format when it pays (F % 128 == 0 and the greedy groups are dense
        enough): the tensor-core format (64-row groups) first, else the CUDA-core RBU format.
        mode: "auto" | "off" | "tc" | "tc16" | "rbu" | "force4/8/16".  "tc16" (or "auto" with
        SGP_B200_HOP=tc16) builds the fp16x3 / 96-row format INSTEAD of the tf32 one; it serves
        launches that come with a bound on |x| (the encoders' tanh states), anything else falls to
        the CSR kernel."""
        if self.rbu is not None or self.tc is not None or self.tc16 is not None or mode == "off" or F % 128 != 0:
            return
        if mode == "auto" and _HOP_DEFAULT == "tc16":
            mode = "auto16"
        if mode == "auto" and self.csr.nnz < _RBU_MIN_NNZ:
            return
        if mode.startswith("force"):
            self.rbu = ops.rbu_build(self.csr, int(mode[len("force"):]), n_cols=self.n_cols)
            return
        if mode in ("auto16", "tc16") and (F // 128) in (1, 2, 4):
            cand16 = ops.tc16_build(self.csr, n_cols=self.n_cols)
            if cand16.fill >= _TC_MIN_FILL * 0.7 or mode == "tc16":
                self.tc16 = cand16
                return
        if mode in ("auto", "auto16", "tc") and (F // 128) in (1, 2, 4):
            cand = ops.tc_build(self.csr, n_cols=self.n_cols)
            if cand.fill >= _TC_MIN_FILL or mode == "tc":
                self.tc = cand
                return
        for R in (16, 8, 4):
            cand = ops.rbu_build(self.csr, R, n_cols=self.n_cols)
            if cand.fill >= _RBU_MIN_FILL[R]:
                self.rbu = cand
                return

    def apply(self, src: Tensor, dst: Tensor, halo: Optional[Tensor] = None,
              checksum: Optional[Tensor] = None, bound: Optional[float] = None) -> None:
        """dst[t] = S @ src[t] for [T, N, F] device views (dst must not alias src); with `halo`
        [T, n_halo, F] the operator's column ids >= n_split read halo rows.  `checksum` (device
        float64 scalar) += sum(dst): fused into the tensor-core hop's epilogue, a separate
        reduction over dst for the CUDA-core kernels."""
        F = src.size(-1)
        aligned = (F % 128 == 0 and src.data_ptr() % 16 == 0 and dst.data_ptr() % 16 == 0 and
                   all(s % 4 == 0 for s in (*src.stride()[:2], *dst.stride()[:2])))
        if (self.tc16 is not None and bound is not None and math.isfinite(bound) and bound > 0 and aligned and
                (F // 128) in (1, 2, 4)):
            ops.spmm_tc16(self.tc16, src, dst, bound, halo, self.n_split, checksum)      # `bound` >= max|src|
            return
        if self.tc is not None and aligned and (F // 128) in (1, 2, 4):
            ops.spmm_tc(self.tc, src, dst, halo, self.n_split, checksum)
            return
        if self.rbu is not None and aligned:
            ops.spmm_rbu(self.rbu, src, dst, halo, self.n_split)
        else:
            ops.spmm(self.csr, src, dst, halo=halo, n_split=self.n_split)
        if checksum is not None:
            ops.checksum_view(dst[:, :self.num_nodes], checksum)

    def check(self) -> None:
        """Raise if a tensor-core launch of this operator reported a barrier timeout (syncs)."""
        if self.tc is not None:
            ops.tc_check(self.tc)
        if self.tc16 is not None:
            ops.tc_check(self.tc16)

    def out_bound(self, bound: Optional[float]) -> Optional[float]:
        """A bound on |S x| given one on |x| (the operator's inf-norm; None stays None)."""
        if bound is None:
            return None
        if self.tc16 is not None:
            return bound * max(self.tc16.inf_norm, 1e-30) * (1 + 1e-6)
        return None

    def index_select(self, dim: int, index: Tensor) -> "ShiftOperator":
        """``adj.index_select(0, node_index)`` of the reference's mini-batch path
        (lib/datasets/iid_dataset.py:113-115): the operator restricted to the rows in `index`
        (kept in that order, repeats allowed); the result maps N source rows to len(index) rows."""
        if dim != 0:
            raise NotImplementedError("only row selection (dim = 0) is used by the reference")
        dev = self.device
        idx = torch.as_tensor(index, device=dev).to(torch.int64).reshape(-1)
        rp = self.csr.rowptr.to(torch.int64)
        cnt = rp[idx + 1] - rp[idx]
        new_rp = torch.zeros(idx.numel() + 1, dtype=torch.int64, device=dev)
        new_rp[1:] = torch.cumsum(cnt, 0)
        total = int(new_rp[-1])
        take = torch.repeat_interleave(rp[idx] - new_rp[:-1], cnt) + torch.arange(total, device=dev)
        csr = ops.Csr(new_rp.to(torch.int32), self.csr.col[take].contiguous(), self.csr.val[take].contiguous(),
                      int(idx.numel()))
        sub = ShiftOperator(csr, None, n_split=self.n_split, n_cols=self.n_cols or self.num_nodes)
        return sub

    def __matmul__(self, x: Tensor) -> Tensor:
        """``adj @ x`` for x [N, F] or [..., N, F]; result on x's device."""
        dev = self.device
        xd = x.detach().to(device=dev, dtype=torch.float32)
        lead = xd.shape[:-2]
        x3 = xd.reshape(-1, xd.size(-2), xd.size(-1)).contiguous()
        out = torch.empty(x3.size(0), self.num_nodes, x3.size(-1), device=dev)
        self.apply(x3, out, bound=panel_bound(x3) if self.tc16 is not None else None)
        return out.reshape(*lead, *out.shape[-2:]).to(x.device)


class SparseAdj:
    """Minimal stand-in for ``torch_sparse.SparseTensor`` as the reference constructs and passes it
    (``SparseTensor(row=row, col=col, value=w, sparse_sizes=(N, N))``, lib/sgp_preprocessing.py:81-82):
    an UN-normalised adjacency in COO form.  ``preprocess_adj`` / ``sgp_spatial_embedding`` /
    ``sgp_spatial_support`` accept this class and — duck-typed on ``coo()`` + ``sparse_sizes()`` — a
    real ``torch_sparse.SparseTensor`` when that package is installed."""

    def __init__(self, row: Tensor, col: Tensor, value: Optional[Tensor] = None, sparse_sizes=None):
        self.row, self.col, self.value = torch.as_tensor(row), torch.as_tensor(col), value
        n = int(max(int(self.row.max()), int(self.col.max())) + 1) if self.row.numel() else 0
        self._sizes = tuple(sparse_sizes) if sparse_sizes is not None else (n, n)

    def coo(self):
        return self.row, self.col, self.value

    def sparse_sizes(self):
        return self._sizes

    def t(self) -> "SparseAdj":
        return SparseAdj(self.col, self.row, self.value, self._sizes[::-1])


def _is_sparse_adj(obj) -> bool:
    return not isinstance(obj, (Tensor, np.ndarray)) and hasattr(obj, "coo") and hasattr(obj, "sparse_sizes")


def _sparse_to_edges(adj):
    """A SparseTensor-like adjacency (entry (i, j) = edge j -> i) back to the reference's edge-list
    convention ``edge_index = [col (source j); row (target i)]`` (``col, row = edge_index``, :80)."""
    row, col, value = adj.coo()
    return torch.stack([torch.as_tensor(col), torch.as_tensor(row)]).to(torch.int64), value, int(adj.sparse_sizes()[0])


def panel_bound(x: Tensor) -> float:
    """max |x| of a device panel (one reduction + a sync): the bound the fp16x3 hop needs when the
    caller has no analytic one."""
    return float(x.abs().max()) if x.numel() else 0.0      # nan / inf / 0 make apply() take the plain CSR kernel


Adj = Union[Tensor, np.ndarray, ShiftOperator, SparseAdj]


def _edges_to_device(edge_index, edge_weight, device):
    if isinstance(edge_index, np.ndarray):                       # sgp_preprocessing.py:73-76
        edge_index = torch.from_numpy(edge_index)
        if edge_weight is not None and isinstance(edge_weight, np.ndarray):
            edge_weight = torch.from_numpy(edge_weight)
    if not isinstance(edge_index, Tensor):
        raise RuntimeError("Edge index must be (edge_index, edge_weight) tuple "
                           "or SparseTensor.")
    ei = edge_index.to(device=device, dtype=torch.int64)
    ew = None if edge_weight is None else torch.as_tensor(edge_weight).to(device=device,
                                                                          dtype=torch.float32)
    return ei, ew


def build_operator(edge_index, edge_weight, num_nodes: int, *, gcn_norm=False, set_diag=False,
                   remove_diag=False, symmetrize=False, transpose=False, normalize=True,
                   device=None) -> ShiftOperator:
    if device is None:
        device = _cuda_device_for(edge_index if isinstance(edge_index, Tensor) else torch.empty(0))
    ei, ew = _edges_to_device(edge_index, edge_weight, device)
    flags = ((_lib.CSR_SET_DIAG if set_diag else 0) | (_lib.CSR_REMOVE_DIAG if remove_diag else 0) |
             (_lib.CSR_GCN_NORM if gcn_norm else 0) | (_lib.CSR_SYMMETRIZE if symmetrize else 0) |
             (_lib.CSR_TRANSPOSE if transpose else 0) | (0 if normalize else _lib.CSR_NO_NORM))
    return ShiftOperator(ops.csr_build(ei, ew, int(num_nodes), flags))


def preprocess_adj(edge_index: Adj, edge_weight: Optional[Tensor] = None,
                   num_nodes: Optional[int] = None, gcn_norm: bool = False, set_diag: bool = True,
                   remove_diag: bool = False) -> ShiftOperator:
    """Reference: lib/sgp_preprocessing.py:67-105.  An already-built :class:`ShiftOperator` is
    passed through (the counterpart of handing the reference a SparseTensor, which it would
    re-normalise; ours is normalised at construction)."""
    if isinstance(edge_index, ShiftOperator):
        return edge_index
    if _is_sparse_adj(edge_index):                               # the SparseTensor branch (:83-84)
        edge_index, edge_weight, n = _sparse_to_edges(edge_index)
        num_nodes = n if num_nodes is None else num_nodes
    if not isinstance(edge_index, (Tensor, np.ndarray)):
        raise RuntimeError("Edge index must be (edge_index, edge_weight) tuple "
                           "or SparseTensor.")
    if num_nodes is None:
        ei = torch.as_tensor(edge_index)
        num_nodes = int(ei.max()) + 1 if ei.numel() else 0
    return build_operator(edge_index, edge_weight, num_nodes, gcn_norm=gcn_norm, set_diag=set_diag,
                          remove_diag=remove_diag)


def _dropout_edges(edge_index, edge_weight, p: float):
    """torch_geometric.utils.dropout_adj (PyG 2.0) with training=True: identity at p == 0,
    otherwise keep each edge with probability 1 - p."""
    if p < 0. or p > 1.:
        raise ValueError(f'Dropout probability has to be between 0 and 1 (got {p}')
    if p == 0.0:
        return edge_index, edge_weight
    ei = torch.as_tensor(edge_index)
    keep = torch.bernoulli(torch.full((ei.size(1),), 1 - p, device=ei.device)).to(torch.bool)
    ew = None if edge_weight is None else torch.as_tensor(edge_weight)[keep]
    return ei[:, keep], ew


def spatial_blocks(k: int, bidirectional: bool) -> int:
    return 1 + k * (2 if bidirectional else 1)


def propagate_into(buf: Tensor, F: int, k: int, fwd: ShiftOperator,
                   bwd: Optional[ShiftOperator], checksum: Optional[Tensor] = None,
                   bound: Optional[float] = None) -> None:
    """buf [T, N, >= blocks*F] on the device with block 0 filled: write S^h x into block h for
    h = 1..k and, with `bwd`, the k hops of the reversed operator into blocks k+1..2k
    (the order of the reference's ``res`` list, sgp_preprocessing.py:200-217)."""
    # `bound` >= max|block 0| (1 for tanh reservoir states) lets the fp16x3 hop pick its panel scale; every
    # hop multiplies it by the operator's inf-norm
    b = bound
    for h in range(1, k + 1):
        fwd.apply(buf[..., (h - 1) * F:h * F], buf[..., h * F:(h + 1) * F], checksum=checksum, bound=b)
        b = fwd.out_bound(b)
    if bwd is not None:
        b = bound
        for h in range(1, k + 1):
            src = buf[..., :F] if h == 1 else buf[..., (k + h - 1) * F:(k + h) * F]
            bwd.apply(src, buf[..., (k + h) * F:(k + h + 1) * F], checksum=checksum, bound=b)
            b = bwd.out_bound(b)


def make_operators(edge_index, edge_weight, num_nodes, *, undirected, add_self_loops,
                   remove_self_loops, bidirectional, device, F: int = 0, rbu: str = "auto"):
    """Forward (and reversed) operators exactly as sgp_spatial_embedding builds them (:182-192,
    :205-216)."""
    if undirected:
        assert bidirectional is False
    if _is_sparse_adj(edge_index):                               # un-normalised SparseTensor-like input
        edge_index, edge_weight, _ = _sparse_to_edges(edge_index)
    if isinstance(edge_index, ShiftOperator):
        fwd, bwd = edge_index, None
        if bidirectional:
            raise SgpError("bidirectional propagation needs the edge list, not a built operator")
    else:
        fwd = build_operator(edge_index, edge_weight, num_nodes, gcn_norm=undirected,
                             set_diag=add_self_loops, remove_diag=remove_self_loops,
                             symmetrize=undirected, device=device)
        bwd = None
        if bidirectional:
            bwd = build_operator(edge_index, edge_weight, num_nodes, gcn_norm=False,
                                 set_diag=add_self_loops, remove_diag=remove_self_loops,
                                 transpose=True, device=device)
    for op in (fwd, bwd):
        if op is not None and F:
            op.maybe_build_rbu(F, rbu)
    return fwd, bwd


class OperatorChain:
    """A product of shift operators kept factored: ``(S_n .. S_1) @ x`` applies S_1 first.  Stands
    in for the sparse-sparse product ``adj_0 @ adj_0`` of sgp_spatial_support (:143-145): the K2
    kernels apply the factors one after the other instead of materialising S^2 (same result up to
    fp32 rounding, no fill-in)."""

    def __init__(self, factors):
        self.factors = list(factors)             # applied left to right: factors[0] first

    @property
    def device(self):
        return self.factors[0].device

    def sparse_sizes(self):
        return (self.factors[-1].num_nodes, self.factors[0].sparse_sizes()[1])

    def __matmul__(self, x: Tensor) -> Tensor:
        y = x.detach().to(device=self.device, dtype=torch.float32)
        for f in self.factors:
            y = f @ y
        return y.to(x.device)

    def index_select(self, dim: int, index: Tensor) -> "OperatorChain":
        return OperatorChain(self.factors[:-1] + [self.factors[-1].index_select(dim, index)])


class MeanOperator:
    """The dense ``torch.full((N, N), 1 / N)`` support of ``global_attr`` (:155-158), never
    materialised: ``@ x`` is the node mean broadcast to every output row (kernel K4)."""

    def __init__(self, num_nodes: int, n_rows: Optional[int] = None):
        self.num_nodes, self.n_rows = int(num_nodes), int(num_nodes if n_rows is None else n_rows)

    def sparse_sizes(self):
        return (self.n_rows, self.num_nodes)

    def __matmul__(self, x: Tensor) -> Tensor:
        dev = _cuda_device_for(x)
        xd = x.detach().to(device=dev, dtype=torch.float32)
        lead = xd.shape[:-2]
        x3 = xd.reshape(-1, xd.size(-2), xd.size(-1)).contiguous()
        sums = torch.empty(x3.size(0), x3.size(-1), device=dev)
        ops.node_sum(x3, sums)
        out = torch.empty(x3.size(0), self.n_rows, x3.size(-1), device=dev)
        ops.node_mean_broadcast(sums, self.num_nodes, out)
        return out.reshape(*lead, *out.shape[-2:]).to(x.device)

    def index_select(self, dim: int, index: Tensor) -> "MeanOperator":
        return MeanOperator(self.num_nodes, int(torch.as_tensor(index).numel()))


def sgp_spatial_support(edge_index: Adj, edge_weight=None, num_nodes=None, k=2, undirected=False,
                        add_self_loops=False, remove_self_loops=False, bidirectional=False,
                        global_attr=False) -> list:
    """Reference: lib/sgp_preprocessing.py:108-160 — the operators the on-the-fly models apply to
    mini-batches (``SGPLoader.collate``, ``IIDDataset._populate_input_frame``).  Returns objects
    supporting ``op @ x`` and ``op.index_select(0, node_index)`` where the reference returns
    torch_sparse SparseTensors / a dense matrix.  Reference behaviour kept on purpose:

    * the list is ``[S, S^2, S^2, ...]``: every entry after the first is ``adj_0 @ adj_0`` (:143-145),
      never a higher power;
    * ``bidirectional`` recurses on the assembled (un-normalised, un-transposed) adjacency with
      every flag off (:147-154), so the "backward" operators are the row-normalised forward
      operator again (identical to the forward list unless ``undirected`` changed the
      normalisation);
    * ``global_attr`` appends the N x N matrix of 1/N (:155-158), here an implicit mean operator.
    """
    if isinstance(edge_index, ShiftOperator):
        raise SgpError("sgp_spatial_support needs the edge list: a built ShiftOperator is already normalised")
    if _is_sparse_adj(edge_index):
        edge_index, edge_weight, n = _sparse_to_edges(edge_index)
        num_nodes = n if num_nodes is None else num_nodes
    if num_nodes is None:
        ei = torch.as_tensor(edge_index)
        num_nodes = int(ei.max()) + 1 if ei.numel() else 0
    dev = _cuda_device_for(edge_index if isinstance(edge_index, Tensor) else torch.empty(0))

    def powers(op):
        return [op] + [OperatorChain([op, op]) for _ in range(k - 1)]

    adj_0 = build_operator(edge_index, edge_weight, num_nodes, gcn_norm=undirected, set_diag=add_self_loops,
                           remove_diag=remove_self_loops, symmetrize=undirected, device=dev)
    support = powers(adj_0)
    if bidirectional:
        back = adj_0 if not undirected else build_operator(
            edge_index, edge_weight, num_nodes, gcn_norm=False, set_diag=add_self_loops,
            remove_diag=remove_self_loops, symmetrize=True, device=dev)
        support += powers(back)
    if global_attr:
        support.append(MeanOperator(num_nodes))
    return support


def sgp_collate_features(x: Tensor, support: list, node_index: Optional[Tensor] = None) -> Tensor:
    """The feature assembly of ``SGPLoader.collate`` (lib/dataloader/sgp_dataloader.py:61-66):
    ``cat([x] + [adj @ x for adj in support], -1)``; with ``node_index`` the row-subset form of
    ``IIDDataset._populate_input_frame`` (lib/datasets/iid_dataset.py:112-116):
    ``cat([x.index_select(-2, idx)] + [adj.index_select(0, idx) @ x ...], -1)``."""
    if node_index is None:
        return torch.cat([x] + [(adj @ x).to(x.device) for adj in support], dim=-1)
    idx = torch.as_tensor(node_index).to(torch.int64).reshape(-1)
    return torch.cat([x.index_select(-2, idx.to(x.device))] +
                     [(adj.index_select(0, idx) @ x).to(x.device) for adj in support], dim=-1)


def round_chunk_steps(steps: int, n_steps: int) -> int:
    """Chunks of a multiple of 4 time steps: the tensor-core hop works on time blocks of 4 / (F / 128)
    steps per work item, and a ragged last block costs a whole one (5 steps = 2 blocks of 4 at F = 128)."""
    steps = max(1, min(int(steps), int(n_steps)))
    return steps - steps % 4 if 4 <= steps < n_steps else steps


def _chunk_steps(n_steps: int, bytes_per_step: int) -> int:
    budget = int(os.environ.get("SGP_B200_CHUNK_BYTES", 4 << 30))
    return round_chunk_steps(budget // max(bytes_per_step, 1), n_steps)


def sgp_spatial_embedding(x, num_nodes, edge_index, edge_weight=None, k=2, undirected=False,
                          add_self_loops=False, remove_self_loops=False, bidirectional=False,
                          one_hot_encoding=False, dropout_rate=0.) -> List[Tensor]:
    """Reference: lib/sgp_preprocessing.py:163-218.  x [batch, node, features] on CPU or GPU;
    returns ``[x, Sx, .., S^k x (, reversed-graph hops)]`` on x's device as views of one
    concatenated buffer."""
    edge_index, edge_weight = _dropout_edges(edge_index, edge_weight, dropout_rate)
    dev = _cuda_device_for(x)
    squeeze = x.dim() == 2
    x3 = x[None] if squeeze else x
    B, N, F0 = x3.shape
    F = F0 + (num_nodes if one_hot_encoding else 0)
    fwd, bwd = make_operators(edge_index, edge_weight, num_nodes, undirected=undirected,
                              add_self_loops=add_self_loops, remove_self_loops=remove_self_loops,
                              bidirectional=bidirectional, device=dev, F=F)
    nb = spatial_blocks(k, bidirectional)
    out = torch.empty(B, N, nb * F, dtype=torch.float32, device=x.device)
    step = _chunk_steps(B, N * nb * F * 4) if not x.is_cuda else B
    for t0 in range(0, B, step):
        t1 = min(B, t0 + step)
        buf = out[t0:t1] if x.is_cuda else torch.empty(t1 - t0, N, nb * F, device=dev)
        buf[..., :F0] = x3[t0:t1].to(device=dev, dtype=torch.float32)
        if one_hot_encoding:
            buf[..., F0:F] = torch.eye(num_nodes, device=dev)
        propagate_into(buf, F, k, fwd, bwd, bound=panel_bound(buf[..., :F]) if fwd.tc16 is not None else None)
        if not x.is_cuda:
            out[t0:t1] = buf.to(x.device)
    for op in (fwd, bwd):
        if op is not None:
            op.check()
    res = [out[..., b * F:(b + 1) * F] for b in range(nb)]
    return [r[0] for r in res] if squeeze else res


def reservoir_preprocessing_(data, hidden_size: int, preprocess_exogenous=False, num_layers=1,
                             leaking_rate=0.9, spectral_radius=0.9, density=0.9, activation='tanh',
                             bias=True, cuda=False):
    """Reference: lib/sgp_preprocessing.py:40-64.  ``cuda`` is accepted and ignored: the scan always
    runs on the GPU and the result comes back on ``data``'s device, like the reference's
    ``.to(device)``."""
    reservoir = Reservoir(input_size=data.size(-1), hidden_size=hidden_size, num_layers=num_layers,
                          leaking_rate=leaking_rate, spectral_radius=spectral_radius, density=density,
                          activation=activation, bias=bias)
    return reservoir(data[None])[0].to(data.device)


def preprocess_dataset(dataset, preprocess_exogenous, reservoir_kwargs, sgp_kwargs):
    """Reference: lib/sgp_preprocessing.py:15-37 (dataset is a tsl SpatioTemporalDataset or anything
    with the same ``exogenous / get_tensors / edge_index / edge_weight / add_exogenous /
    set_input_map`` surface)."""
    if isinstance(preprocess_exogenous, bool):
        preprocess_exogenous = dataset.exogenous.keys() if preprocess_exogenous else []
    if not isinstance(preprocess_exogenous, (list, tuple)):
        preprocess_exogenous = list(preprocess_exogenous) \
            if not isinstance(preprocess_exogenous, str) else [preprocess_exogenous]
    data, _ = dataset.get_tensors(['data'] + list(preprocess_exogenous), preprocess=True, cat_dim=-1)
    res = reservoir_preprocessing_(data, **reservoir_kwargs)
    res = sgp_spatial_embedding(res, num_nodes=data.size(1), edge_index=dataset.edge_index,
                                edge_weight=dataset.edge_weight, **sgp_kwargs)
    dataset.add_exogenous('processed_x', torch.cat(res, -1), add_to_input_map=False)
    dataset.set_input_map({'x': ['processed_x']})
