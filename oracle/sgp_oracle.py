"""CPU oracle for the SGP training-free encoder hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``sgp_b200/`` imports this module; only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` do, and there only as the checker / the CPU arm that is timed
*beside* the CUDA path.

What it restates (all citations relative to the reference repository root):

* reservoir weight generation      lib/nn/reservoir/reservoir.py:54-75
* per-layer leak schedule          lib/nn/reservoir/reservoir.py:109-123
* leaky-ESN step and time loop     lib/nn/reservoir/reservoir.py:77-81, 158-186
* adjacency construction/normalise lib/sgp_preprocessing.py:67-105
* K-hop propagation, bidirectional lib/sgp_preprocessing.py:163-218
* global block + concatenation     lib/nn/encoders/sgp_spatial_encoder.py:22-35
* encoder composition              lib/nn/encoders/sgp_encoder.py:45-51,
                                   lib/nn/encoders/sgp_temporal_encoder.py:29-34

Parity pins
-----------
* Reservoir (weights and recurrence): PINNED.  ``tests/golden/reservoir_*.npz`` were produced
  by executing the *unmodified* reference file ``lib/nn/reservoir/reservoir.py`` (see
  ``tests/golden/make_golden.py``); ``tests/test_oracle.py`` checks this oracle against them
  bit-for-bit on the weights and to 1e-6 on the states.
* Spatial path: PARITY UNPINNED.  Its arithmetic lives in ``torch_sparse`` (rusty1s/pytorch_sparse,
  pulled transitively by ``pyg=2.0`` in conda_env.yml:8-10, version never pinned, ~0.6.12) and
  ``torch_geometric.utils`` (``dropout_adj``, ``to_undirected``); neither is vendored, installed
  or installable here, and the reference ships no test or golden vector for this boundary.
  The restatement follows the published semantics of those libraries (noted inline) and is
  anchored on hand-computed known-answer cases plus a dense float64 ``matrix_power`` cross-check.

Everything here is written functionally (plain arrays in, plain arrays out) on purpose: it is a
second, independent statement of the algorithm, not a mirror of the product's class layout.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))

ACTIVATIONS = ("tanh", "relu", "self_norm", "identity")


# --------------------------------------------------------------------------------------
# reservoir
# --------------------------------------------------------------------------------------
def layer_alphas(leaking_rate: float, num_layers: int, alpha_decay: bool) -> list:
    """Leak per layer.  reservoir.py:109-123: the running value starts at ``leaking_rate`` and,
    when ``alpha_decay`` is on, is replaced by ``np.clip(alpha - 0.1, 0.1, 1.)`` after each layer
    is built (so it turns into a float64 numpy scalar from layer 1 on)."""
    out, a = [], leaking_rate
    for _ in range(num_layers):
        out.append(a)
        if alpha_decay:
            a = np.clip(a - 0.1, 0.1, 1.0)
    return out


def draw_layer_weights(input_size: int, hidden_size: int, spectral_radius: float,
                       density: float, in_scaling: float = 1.0, bias_scale: float = 1.0):
    """One layer's frozen weights, consuming torch's *global CPU generator* in the reference's
    order (reservoir.py:54-75): w_ih ~ U(-1,1)*in_scaling; b_ih ~ U(-1,1)*bias_scale;
    w_hh ~ U(-1,1); if density < 1 zero ``randperm(H*H)[:int(H*H*(1-density))]``; rescale w_hh
    by spectral_radius / max|eig|.  The bias always exists (reservoir.py:47-51 tests
    ``bias is not None``)."""
    H = hidden_size
    w_ih = torch.empty(H, input_size).uniform_(-1, 1).mul_(in_scaling)
    b_ih = torch.empty(H).uniform_(-1, 1).mul_(bias_scale)
    w_hh = torch.empty(H, H).uniform_(-1, 1)
    if density < 1:
        n_units = H * H
        keep = torch.ones(n_units)
        dropped = torch.randperm(n_units)[: int(n_units * (1 - density))]
        keep[dropped] = 0.0
        w_hh.mul_(keep.view(H, H))
    radius = torch.linalg.eigvals(w_hh).abs().max()
    w_hh.mul_(spectral_radius / radius)
    return w_ih, w_hh, b_ih


def draw_reservoir(input_size: int, hidden_size: int, num_layers: int = 1,
                   leaking_rate: float = 0.9, spectral_radius: float = 0.9,
                   density: float = 0.9, input_scaling: float = 1.0,
                   alpha_decay: bool = False) -> List[dict]:
    """All layers, in construction order (reservoir.py:107-125).  Layer 0 reads ``input_size``
    features, deeper layers read the previous layer's H-wide state."""
    alphas = layer_alphas(leaking_rate, num_layers, alpha_decay)
    layers = []
    for i in range(num_layers):
        w_ih, w_hh, b_ih = draw_layer_weights(input_size if i == 0 else hidden_size, hidden_size,
                                              spectral_radius, density, in_scaling=input_scaling)
        layers.append(dict(w_ih=w_ih, w_hh=w_hh, b_ih=b_ih, alpha=alphas[i]))
    return layers


def _activate(z: torch.Tensor, activation: str) -> torch.Tensor:
    if activation == "tanh":
        return torch.tanh(z)
    if activation == "relu":
        return torch.relu(z)
    if activation == "identity":
        return z
    if activation == "self_norm":  # lib/utils.py:50-51 -> r * F.normalize(x, p=2, dim=-1), r = 1
        return z / z.norm(p=2, dim=-1, keepdim=True).clamp_min(1e-12)
    raise AssertionError(f"activation {activation!r} not in {ACTIVATIONS}")


def reservoir_states(x, layers: Sequence[dict], activation: str = "tanh",
                     h0: Optional[torch.Tensor] = None, dtype=torch.float32,
                     return_last: bool = False):
    """Run the stacked leaky ESN over ``x`` [T, N, Fin] -> [T, N, L*H] (layer-major features).

    reservoir.py:158-186 with b = 1: zero initial state unless ``h0`` [L, N, H] is given; at
    each step layer i consumes the *new* state of layer i-1 of the same step; the update is
    ``h' = (1-a) h + a act(x W_ih^T + b + h W_hh^T)`` (reservoir.py:77-81).  ``dtype=float64``
    gives the high-precision variant used for error budgeting."""
    x = torch.as_tensor(x).to(dtype)
    T, N, _ = x.shape
    L, H = len(layers), layers[0]["w_hh"].shape[0]
    W = [(l["w_ih"].to(dtype), l["w_hh"].to(dtype), l["b_ih"].to(dtype)) for l in layers]
    h = torch.zeros(L, N, H, dtype=dtype) if h0 is None else torch.as_tensor(h0).to(dtype).clone()
    out = torch.empty(T, N, L * H, dtype=dtype)
    for t in range(T):
        inp = x[t]
        for i, (w_ih, w_hh, b) in enumerate(W):
            a = layers[i]["alpha"]
            pre = torch.addmm(b, inp, w_ih.t()) + h[i] @ w_hh.t()
            new = (1 - a) * h[i] + a * _activate(pre, activation)
            h[i] = new
            out[t, :, i * H:(i + 1) * H] = new
            inp = new
    if return_last:
        return out, h
    return out


# --------------------------------------------------------------------------------------
# adjacency
# --------------------------------------------------------------------------------------
def undirected_edges(edge_index: np.ndarray, edge_weight: Optional[np.ndarray], num_nodes: int):
    """``torch_geometric.utils.to_undirected`` (PyG 2.0): append every edge reversed (weights
    duplicated), then coalesce duplicates by (first row, second row) with *add*; the result is
    sorted by that key."""
    a, b = np.asarray(edge_index[0], np.int64), np.asarray(edge_index[1], np.int64)
    r = np.concatenate([a, b])
    c = np.concatenate([b, a])
    key = r * num_nodes + c
    uniq, inv = np.unique(key, return_inverse=True)
    ei = np.stack([uniq // num_nodes, uniq % num_nodes])
    if edge_weight is None:
        return ei, None
    w = np.concatenate([edge_weight, edge_weight]).astype(np.float32)
    acc = np.zeros(uniq.shape[0], np.float32)
    np.add.at(acc, inv, w)
    return ei, acc


def build_operator(edge_index, edge_weight, num_nodes: int, gcn_norm: bool = False,
                   set_diag: bool = True, remove_diag: bool = False):
    """CSR (rowptr int64 [N+1], col int64 [nnz], val float32 [nnz]) of the normalised shift
    operator.  lib/sgp_preprocessing.py:67-105 on top of torch_sparse semantics:

    * ``col, row = edge_index`` (:80): edge_index[0] is the *column* (source), [1] the row.
    * SparseTensor orders entries by (row, col) and keeps duplicates.
    * ``set_diag()`` drops every stored diagonal entry and inserts value 1 at (i, i) for all i,
      in sorted position; ``remove_diag()`` only drops (:89-92; set_diag wins when both are set).
    * ``sum(dim=1)`` = row sums of the values (row counts when there are no values).
    * row normalisation ``D^-1 S`` with ``inf -> 0`` (:100-103) or ``D^-1/2 S D^-1/2`` (:95-98),
      evaluated as ``(d[row] * v) * d[col]`` in float32.
    """
    if not isinstance(edge_index, (np.ndarray, torch.Tensor)):
        raise RuntimeError("Edge index must be (edge_index, edge_weight) tuple or SparseTensor.")
    ei = np.asarray(edge_index, dtype=np.int64)
    col, row = ei[0], ei[1]
    N = int(num_nodes)
    w = (np.ones(row.shape[0], np.float32) if edge_weight is None
         else np.asarray(edge_weight, dtype=np.float32))
    if set_diag or remove_diag:
        off = row != col
        row, col, w = row[off], col[off], w[off]
    if set_diag:
        d = np.arange(N, dtype=np.int64)
        row, col = np.concatenate([row, d]), np.concatenate([col, d])
        w = np.concatenate([w, np.ones(N, np.float32)])
    order = np.argsort(row * N + col, kind="stable")
    row, col, w = row[order], col[order], w[order]
    rowptr = np.zeros(N + 1, np.int64)
    np.cumsum(np.bincount(row, minlength=N), out=rowptr[1:])
    deg = np.zeros(N, np.float32)
    np.add.at(deg, row, w)
    with np.errstate(divide="ignore"):
        if gcn_norm:
            s = np.power(deg, np.float32(-0.5), dtype=np.float32)
            s[np.isinf(s)] = 0
            val = (s[row] * w) * s[col]
        else:
            s = np.power(deg, np.float32(-1.0), dtype=np.float32)
            s[np.isinf(s)] = 0
            val = s[row] * w
    return rowptr, col.astype(np.int64), val.astype(np.float32)


def csr_to_dense(rowptr, col, val, num_nodes: int, dtype=np.float64) -> np.ndarray:
    A = np.zeros((num_nodes, num_nodes), dtype)
    rows = np.repeat(np.arange(num_nodes), np.diff(rowptr))
    np.add.at(A, (rows, col), val.astype(dtype))
    return A


# --------------------------------------------------------------------------------------
# CSR x dense  (torch_sparse spmm_sum restated)
# --------------------------------------------------------------------------------------
def spmm_loops(rowptr, col, val, x: np.ndarray) -> np.ndarray:
    """Pure-python/numpy statement of torch_sparse ``spmm_cpu`` with reduce=sum: for every
    (batch b, row m) accumulate ``val[e] * x[b, col[e], :]`` over the row's entries in stored
    order, in the dtype of ``x``.  Small inputs only."""
    B, N, F = x.shape
    out = np.zeros((B, len(rowptr) - 1, F), x.dtype)
    for b in range(B):
        for m in range(len(rowptr) - 1):
            acc = np.zeros(F, x.dtype)
            for e in range(rowptr[m], rowptr[m + 1]):
                acc = acc + val[e].astype(x.dtype) * x[b, col[e]]
            out[b, m] = acc
    return out


_clib = None


def build_c(force: bool = False) -> str:
    """Compile oracle/sgp_oracle.c into oracle/_build/libsgp_oracle.so (gcc, OpenMP)."""
    out_dir = os.path.join(_HERE, "_build")
    so = os.path.join(out_dir, "libsgp_oracle.so")
    src = os.path.join(_HERE, "sgp_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-mavx2", "-mfma", "-fopenmp", "-fPIC", "-shared",
                               "-o", so, src, "-lm"])
    return so


def _c():
    global _clib
    if _clib is None:
        lib = ctypes.CDLL(build_c())
        lib.oracle_spmm_csr_f32.restype = None
        lib.oracle_spmm_csr_f32.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int64] * 3 + \
            [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
             ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64]
        lib.oracle_num_threads.restype = ctypes.c_int
        _clib = lib
    return _clib


def c_threads() -> int:
    return int(_c().oracle_num_threads())


def spmm_c(rowptr, col, val, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The timed CPU SpMM: C/OpenMP restatement (oracle/sgp_oracle.c) of torch_sparse's loop
    order.  ``x`` and ``out`` are float32 [B, N, F] views whose last dim is contiguous (row and
    batch strides are passed through, so slices of the concatenated output work)."""
    rowptr = np.ascontiguousarray(rowptr, np.int64)
    col = np.ascontiguousarray(col, np.int64)
    val = np.ascontiguousarray(val, np.float32)
    assert x.dtype == torch.float32 and x.stride(-1) == 1
    B, N, F = x.shape
    M = len(rowptr) - 1
    if out is None:
        out = torch.empty(B, M, F, dtype=torch.float32)
    assert out.stride(-1) == 1
    _c().oracle_spmm_csr_f32(rowptr.ctypes.data, col.ctypes.data, val.ctypes.data, B, M, F,
                             x.data_ptr(), x.stride(0), x.stride(1),
                             out.data_ptr(), out.stride(0), out.stride(1))
    return out


def spmm(rowptr, col, val, x, impl: str = "scipy"):
    """CSR x dense over the batch (time) axis.  impl: 'scipy' (float of x), 'loops', 'c'."""
    if impl == "c":
        return spmm_c(rowptr, col, val, torch.as_tensor(x)).numpy()
    x = np.asarray(x)
    if impl == "loops":
        return spmm_loops(rowptr, col, val, x)
    import scipy.sparse as sp
    B, N, F = x.shape
    M = len(rowptr) - 1
    A = sp.csr_matrix((val.astype(x.dtype), col, rowptr), shape=(M, N))
    y = A @ np.ascontiguousarray(x.transpose(1, 0, 2)).reshape(N, B * F)
    return np.ascontiguousarray(y.reshape(M, B, F).transpose(1, 0, 2))


# --------------------------------------------------------------------------------------
# spatial embedding / encoders
# --------------------------------------------------------------------------------------
def spatial_embedding(x, num_nodes: int, edge_index, edge_weight=None, k: int = 2,
                      undirected: bool = False, add_self_loops: bool = False,
                      remove_self_loops: bool = False, bidirectional: bool = False,
                      one_hot_encoding: bool = False, impl: str = "scipy") -> List[np.ndarray]:
    """lib/sgp_preprocessing.py:163-218 (dropout_rate = 0, the only value any caller passes, is
    the identity in PyG's dropout_adj).  Returns ``[x, Sx, ..., S^k x]`` and, when bidirectional,
    the k hops of the independently normalised reversed graph appended (:205-217)."""
    x = np.asarray(x)
    ei = np.asarray(edge_index, np.int64)
    ew = None if edge_weight is None else np.asarray(edge_weight, np.float32)
    fwd_ei, fwd_ew = ei, ew
    if undirected:
        assert bidirectional is False
        fwd_ei, fwd_ew = undirected_edges(ei, ew, num_nodes)
    rowptr, col, val = build_operator(fwd_ei, fwd_ew, num_nodes, gcn_norm=undirected,
                                      set_diag=add_self_loops, remove_diag=remove_self_loops)
    if one_hot_encoding:
        ids = np.broadcast_to(np.eye(num_nodes, dtype=x.dtype), (x.shape[0], num_nodes, num_nodes))
        x = np.concatenate([x, ids], -1)
    res = [x]
    for _ in range(k):
        res.append(spmm(rowptr, col, val, res[-1], impl=impl))
    if bidirectional:
        # the recursive call receives the (possibly symmetrised) edge list with rows swapped
        back = spatial_embedding(res[0], num_nodes, fwd_ei[[1, 0]], fwd_ew, k=k, undirected=False,
                                 add_self_loops=add_self_loops,
                                 remove_self_loops=remove_self_loops, bidirectional=False,
                                 one_hot_encoding=False, impl=impl)
        res += back[1:]
    return res


def spatial_encoder(x, edge_index, edge_weight, receptive_field: int, bidirectional: bool,
                    undirected: bool, global_attr: bool, add_self_loops: bool = False,
                    impl: str = "scipy") -> np.ndarray:
    """lib/nn/encoders/sgp_spatial_encoder.py:22-35: blocks ``[x | Sx | .. | S^K x | reversed
    hops | node-mean]`` concatenated on the feature axis."""
    x = np.asarray(x)
    out = spatial_embedding(x, x.shape[-2], edge_index, edge_weight, k=receptive_field,
                            bidirectional=bidirectional, undirected=undirected,
                            add_self_loops=add_self_loops, impl=impl)
    if global_attr:
        out.append(np.ones_like(x) * x.mean(-2, keepdims=True))
    return np.concatenate(out, -1)


def sgp_encoder(x, edge_index, edge_weight, layers: Sequence[dict], activation: str,
                receptive_field: int, bidirectional: bool, undirected: bool, global_attr: bool,
                add_self_loops: bool = False, impl: str = "scipy",
                dtype=torch.float32) -> np.ndarray:
    """lib/nn/encoders/sgp_encoder.py:45-51: reservoir over [T,N,Fin], then the spatial encoder."""
    h = reservoir_states(x, layers, activation, dtype=dtype).numpy()
    if impl == "c":
        h = h.astype(np.float32, copy=False)       # the C SpMM is float32 (a float64 recurrence is rounded once)
    return spatial_encoder(h, edge_index, edge_weight, receptive_field, bidirectional,
                           undirected, global_attr, add_self_loops, impl=impl)


def blockwise_allclose(got, ref, block: int, rtol: float = 1e-4, atol_rel: float = 1e-5):
    """The parity metric of SURVEY.md 8(c): per feature block of width ``block``,
    |got - ref| <= atol_rel * max|ref_block| + rtol * |ref|.  Returns (ok, worst_ratio)."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    worst = 0.0
    for s in range(0, ref.shape[-1], block):
        r, g = ref[..., s:s + block], got[..., s:s + block]
        if r.size == 0:
            continue
        tol = atol_rel * np.abs(r).max() + rtol * np.abs(r)
        tol = np.maximum(tol, 1e-30)
        worst = max(worst, float((np.abs(g - r) / tol).max()))
    return worst <= 1.0, worst


# --------------------------------------------------------------------------------------
# the callers either side of the hot path (restated for the parity tests)
# --------------------------------------------------------------------------------------
def iid_sample(x: np.ndarray, y: np.ndarray, step_index: np.ndarray, node_index: np.ndarray,
               horizon: int, delay: int = 0, horizon_lag: int = 1):
    """``IIDDataset.sample`` (lib/datasets/iid_dataset.py:57-99) for given indices:
    ``tens[(step_index, None, None, node_index)]`` -> [B, 1, 1, D] and
    ``tens[(hor_index, node_index[:, None], None)]`` -> [B, h, 1, C] with
    ``hor_index = stack([step_index + i for i in range(delay + 1, horizon + 1, horizon_lag)], 1)``."""
    xs = x[step_index, None, None, node_index]
    hor = np.stack([step_index + i for i in range(delay + 1, horizon + 1, horizon_lag)], 1)
    ys = y[hor, node_index[:, None], None]
    return xs, ys


def spatial_support_dense(edge_index, edge_weight, num_nodes: int, k: int = 2, undirected: bool = False,
                          add_self_loops: bool = False, remove_self_loops: bool = False,
                          bidirectional: bool = False, global_attr: bool = False, _adj=None) -> List[np.ndarray]:
    """``sgp_spatial_support`` (lib/sgp_preprocessing.py:108-160) with dense float64 matrices, line
    by line — including ``support.append(adj_0 @ adj_0)`` for EVERY extra order (:143-145), the
    recursion on the assembled, un-transposed ``adj`` for ``bidirectional`` (:147-154) and the dense
    1/N matrix for ``global_attr`` (:155-158)."""
    N = int(num_nodes)
    if _adj is None:
        ei = np.asarray(edge_index, np.int64)
        col, row = ei[0], ei[1]
        A = np.zeros((N, N))
        w = np.ones(ei.shape[1]) if edge_weight is None else np.asarray(edge_weight, np.float64)
        np.add.at(A, (row, col), w)                      # SparseTensor(row, col, value): duplicates summed by @
    else:
        A = _adj
    if undirected:
        A = A + A.T
    if add_self_loops:
        A = A.copy()
        np.fill_diagonal(A, 1.0)
    elif remove_self_loops:
        A = A.copy()
        np.fill_diagonal(A, 0.0)
    deg = A.sum(1)
    with np.errstate(divide="ignore"):
        if undirected:
            d = deg ** -0.5
            d[np.isinf(d)] = 0
            A0 = d[:, None] * A * d[None, :]
        else:
            d = deg ** -1.0
            d[np.isinf(d)] = 0
            A0 = d[:, None] * A
    support = [A0]
    for _ in range(k - 1):
        support.append(A0 @ A0)
    if bidirectional:
        support += spatial_support_dense(None, None, N, k=k, _adj=A)
    if global_attr:
        support.append(np.full((N, N), 1.0 / N))
    return support


def grouped_conv1x1(x: np.ndarray, weight: np.ndarray, bias, groups: int) -> np.ndarray:
    """The reference's own op for SGPModel.input_encoder (lib/nn/models/sgp_model.py:41-52) on the
    CPU, in float64: 'b n f -> b f n', conv1d(kernel_size=1, groups), 'b f n -> b n f'."""
    xt = torch.as_tensor(x, dtype=torch.float64).permute(0, 2, 1)
    y = torch.nn.functional.conv1d(xt, torch.as_tensor(weight, dtype=torch.float64),
                                   None if bias is None else torch.as_tensor(bias, dtype=torch.float64),
                                   groups=groups)
    return y.permute(0, 2, 1).numpy()


def gesn_operator_dense(edge_index, edge_weight, num_nodes: int) -> np.ndarray:
    """``GESNEncoder.forward`` (lib/nn/encoders/dyn_gesn_encoder.py:34-41): PyG ``add_self_loops``
    (unit loops appended for nodes 0..edge_index.max(), stored diagonals kept), tsl
    ``normalize(dim=1)`` (divide by the weighted degree of ``edge_index[1]``), then
    ``col, row = edge_index``."""
    ei = np.asarray(edge_index, np.int64)
    w = np.asarray(edge_weight, np.float64)
    n_loops = int(ei.max()) + 1 if ei.size else 0
    loops = np.arange(n_loops)
    ei = np.concatenate([ei, np.stack([loops, loops])], 1)
    w = np.concatenate([w, np.ones(n_loops)])
    deg = np.zeros(num_nodes)
    np.add.at(deg, ei[1], w)
    w = w / deg[ei[1]]
    S = np.zeros((num_nodes, num_nodes))
    np.add.at(S, (ei[1], ei[0]), w)
    return S


def draw_graph_esn(input_size: int, hidden_size: int, num_layers: int = 1, leaking_rate: float = 0.9,
                   spectral_radius: float = 0.9, density: float = 0.9, input_scaling: float = 1.0,
                   alpha_decay: bool = False) -> List[dict]:
    """``GraphESN.__init__`` (lib/nn/reservoir/graph_reservoir.py:96-144): every GESNLayer draws in its
    constructor, then ``self.reset_parameters()`` draws all layers again; the second draw stays."""
    draw_reservoir(input_size, hidden_size, num_layers, leaking_rate, spectral_radius, density,
                   input_scaling, alpha_decay)
    return draw_reservoir(input_size, hidden_size, num_layers, leaking_rate, spectral_radius, density,
                          input_scaling, alpha_decay)


def graph_esn_states(x, layers: Sequence[dict], S: np.ndarray, activation: str = "tanh") -> np.ndarray:
    """``GraphESN`` over time in float64 (graph_reservoir.py:85-93 inside tsl's _GraphRNN loop,
    tsl/nn/blocks/encoders/gcrnn.py:57-93): per step and layer
    ``h' = (1 - a) h + a act(W_ih x + b + S (h W_hh^T))``; layer l > 0 reads layer l-1's NEW state;
    the output concatenates all layers' states."""
    x = torch.as_tensor(np.asarray(x), dtype=torch.float64)
    St = torch.as_tensor(S, dtype=torch.float64)
    T, N, _ = x.shape
    H = layers[0]["w_hh"].shape[0]
    h = [torch.zeros(N, H, dtype=torch.float64) for _ in layers]
    out = torch.empty(T, N, len(layers) * H, dtype=torch.float64)
    for t in range(T):
        inp = x[t]
        for i, l in enumerate(layers):
            pre = inp @ l["w_ih"].double().t() + l["b_ih"].double() + St @ (h[i] @ l["w_hh"].double().t())
            h[i] = (1 - float(l["alpha"])) * h[i] + float(l["alpha"]) * _activate(pre, activation)
            inp = h[i]
            out[t, :, i * H:(i + 1) * H] = h[i]
    return out.numpy()
