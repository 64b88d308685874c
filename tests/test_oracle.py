"""The oracle against the reference's own outputs (golden .npz made by tests/golden/make_golden.py
from the unmodified reference reservoir.py) and against hand-computed known answers for the
spatial path."""
import numpy as np
import pytest
import torch

from oracle import sgp_oracle as O
from tests.helpers import golden_names, load_golden, random_graph


@pytest.mark.parametrize("name", golden_names())
def test_weights_bit_exact_vs_reference(name):
    g = load_golden(name)
    kw = dict(g["kwargs"])
    torch.manual_seed(g["seed"])
    layers = O.draw_reservoir(kw["input_size"], kw["hidden_size"], kw.get("num_layers", 1),
                              kw.get("leaking_rate", 0.9), kw.get("spectral_radius", 0.9),
                              kw.get("density", 0.9), kw.get("input_scaling", 1.0),
                              kw.get("alpha_decay", False))
    for got, ref in zip(layers, g["layers"]):
        for k in ("w_ih", "w_hh", "b_ih"):
            assert torch.equal(got[k], ref[k]), k
        assert float(got["alpha"]) == ref["alpha"]


@pytest.mark.parametrize("name", golden_names())
def test_states_vs_reference(name):
    g = load_golden(name)
    y = O.reservoir_states(g["x"], g["layers"], g["kwargs"].get("activation", "tanh")).numpy()
    np.testing.assert_allclose(y, g["y"], rtol=1e-5, atol=1e-6)
    y64 = O.reservoir_states(g["x"], g["layers"], g["kwargs"].get("activation", "tanh"),
                             dtype=torch.float64).numpy()
    np.testing.assert_allclose(y64, g["y"], rtol=1e-4, atol=2e-5)


def test_alpha_schedule():
    assert O.layer_alphas(0.9, 3, False) == [0.9, 0.9, 0.9]
    a = O.layer_alphas(0.25, 4, True)
    np.testing.assert_allclose(a, [0.25, 0.15, 0.1, 0.1])


# ---- spatial known answers -----------------------------------------------------------
def path3():
    # dense A (row = target i, col = source j): 0<-1 (w 2), 1<-0 (w 1), 1<-2 (w 3), 2<-1 (w 4)
    A = np.array([[0, 2, 0], [1, 0, 3], [0, 4, 0]], np.float64)
    # reference convention: edge_index = nonzero(A^T) -> [0] = col j, [1] = row i
    jj, ii = np.nonzero(A.T)
    return A, np.stack([jj, ii]).astype(np.int64), A.T[jj, ii].astype(np.float32)


def test_kat_path_graph_row_norm():
    A, ei, ew = path3()
    rowptr, col, val = O.build_operator(ei, ew, 3, set_diag=False)
    S = O.csr_to_dense(rowptr, col, val, 3)
    np.testing.assert_allclose(S, [[0, 1, 0], [0.25, 0, 0.75], [0, 1, 0]], rtol=1e-6)
    assert rowptr.tolist() == [0, 1, 3, 4] and col.tolist() == [1, 0, 2, 1]
    x = np.array([[[1.0], [2.0], [4.0]]], np.float32)
    res = O.spatial_embedding(x, 3, ei, ew, k=2)
    np.testing.assert_allclose(res[1][0, :, 0], [2.0, 3.25, 2.0], rtol=1e-6)
    np.testing.assert_allclose(res[2][0, :, 0], [3.25, 2.0, 3.25], rtol=1e-6)


def test_kat_isolated_node_and_duplicates():
    # node 2 has no incoming edge -> zero row (inf -> 0); edge 1->0 appears twice -> summed
    ei = np.array([[1, 1, 0], [0, 0, 1]], np.int64)
    ew = np.array([1.0, 3.0, 5.0], np.float32)
    rowptr, col, val = O.build_operator(ei, ew, 3, set_diag=False)
    assert rowptr.tolist() == [0, 2, 3, 3] and col.tolist() == [1, 1, 0]
    np.testing.assert_allclose(val, [0.25, 0.75, 1.0])
    x = np.array([[[1.0], [10.0], [100.0]]], np.float32)
    y = O.spatial_embedding(x, 3, ei, ew, k=1)[1]
    np.testing.assert_allclose(y[0, :, 0], [10.0, 1.0, 0.0])


def test_kat_self_loops_replace_diagonal():
    ei = np.array([[0, 1], [0, 0]], np.int64)       # (0<-0, w 7) and (0<-1, w 3)
    ew = np.array([7.0, 3.0], np.float32)
    rowptr, col, val = O.build_operator(ei, ew, 2, set_diag=True)
    S = O.csr_to_dense(rowptr, col, val, 2)
    np.testing.assert_allclose(S, [[0.25, 0.75], [0.0, 1.0]])     # diag 7 replaced by 1
    rowptr, col, val = O.build_operator(ei, ew, 2, set_diag=False, remove_diag=True)
    np.testing.assert_allclose(O.csr_to_dense(rowptr, col, val, 2), [[0, 1.0], [0, 0]])


def test_kat_unweighted_counts():
    ei = np.array([[1, 2, 0], [0, 0, 1]], np.int64)
    rowptr, col, val = O.build_operator(ei, None, 3, set_diag=False)
    np.testing.assert_allclose(val, [0.5, 0.5, 1.0])


def test_kat_undirected_gcn_norm():
    A, ei, ew = path3()
    sym = A + A.T
    d = sym.sum(1)
    want = sym / np.sqrt(d)[:, None] / np.sqrt(d)[None, :]
    uei, uew = O.undirected_edges(ei, ew, 3)
    rowptr, col, val = O.build_operator(uei, uew, 3, gcn_norm=True, set_diag=False)
    np.testing.assert_allclose(O.csr_to_dense(rowptr, col, val, 3), want, rtol=1e-6)


def test_bidirectional_is_forward_on_swapped_edges():
    ei, ew = random_graph(11, 40, 3)
    x = np.random.default_rng(0).standard_normal((4, 11, 5)).astype(np.float32)
    res = O.spatial_embedding(x, 11, ei, ew, k=3, bidirectional=True)
    assert len(res) == 7
    back = O.spatial_embedding(x, 11, ei[[1, 0]], ew, k=3)
    for a, b in zip(res[4:], back[1:]):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("impl", ["loops", "scipy", "c"])
def test_khop_equals_dense_matrix_power(impl):
    n = 17
    ei, ew = random_graph(n, 90, 5)
    x = np.random.default_rng(1).standard_normal((3, n, 8)).astype(np.float32)
    rowptr, col, val = O.build_operator(ei, ew, n, set_diag=False)
    S = O.csr_to_dense(rowptr, col, val, n)
    res = O.spatial_embedding(x, n, ei, ew, k=3, impl=impl)
    P = np.eye(n)
    for k in range(4):
        np.testing.assert_allclose(res[k], np.einsum("ij,bjf->bif", P, x.astype(np.float64)),
                                   rtol=1e-4, atol=1e-5)
        P = S @ P


def test_row_stochastic():
    ei, ew = random_graph(50, 300, 7)
    rowptr, col, val = O.build_operator(ei, ew, 50, set_diag=False)
    sums = np.add.reduceat(np.concatenate([val, [0]]), rowptr[:-1])[: 50]
    nz = np.diff(rowptr) > 0
    np.testing.assert_allclose(sums[nz], 1.0, rtol=1e-5)


def test_c_spmm_strided_blocks():
    n, F = 13, 6
    ei, ew = random_graph(n, 60, 9)
    rowptr, col, val = O.build_operator(ei, ew, n, set_diag=False)
    buf = torch.zeros(5, n, 3 * F)
    buf[..., :F] = torch.randn(5, n, F, generator=torch.Generator().manual_seed(0))
    O.spmm_c(rowptr, col, val, buf[..., :F], out=buf[..., F:2 * F])
    want = O.spmm(rowptr, col, val, buf[..., :F].contiguous().numpy(), impl="loops")
    np.testing.assert_allclose(buf[..., F:2 * F].numpy(), want, rtol=1e-5, atol=1e-6)
    assert float(buf[..., 2 * F:].abs().max()) == 0.0


def test_spatial_encoder_layout_and_global_block():
    n, F = 9, 4
    ei, ew = random_graph(n, 30, 2)
    x = np.random.default_rng(2).standard_normal((3, n, F)).astype(np.float32)
    y = O.spatial_encoder(x, ei, ew, receptive_field=2, bidirectional=True, undirected=False,
                          global_attr=True)
    assert y.shape == (3, n, F * 6)
    np.testing.assert_array_equal(y[..., :F], x)
    np.testing.assert_allclose(y[..., 5 * F:], np.broadcast_to(x.mean(1, keepdims=True), x.shape),
                               rtol=1e-6)


def test_encoder_composition_layer_major():
    g = load_golden("tanh_l2_decay")
    n = g["x"].shape[1]
    ei, ew = random_graph(n, 20, 4)
    y = O.sgp_encoder(g["x"], ei, ew, g["layers"], "tanh", receptive_field=1, bidirectional=False,
                      undirected=False, global_attr=False)
    F = g["y"].shape[-1]
    np.testing.assert_allclose(y[..., :F], g["y"], rtol=1e-5, atol=1e-6)   # block 0 = reservoir


def test_bad_edge_index_type():
    with pytest.raises(RuntimeError):
        O.build_operator([[0], [1]], None, 2)


# ---- every flag combination against two independent sparse libraries -----------------------
# torch_sparse / PyG cannot be installed here (parity of the spatial path stays "unpinned"), so the
# restatement of their semantics is cross-checked, for EVERY flag combination the product builds,
# against a dense float64 construction written straight from the reference's source lines and
# against scipy.sparse and torch.sparse_csr matmuls of the oracle's CSR.
ALL_FLAG_CASES = [dict(gcn_norm=g, set_diag=s, remove_diag=r, symmetrize=u, transpose=t)
                  for g in (False, True) for s in (False, True) for r in (False, True)
                  for u in (False, True) for t in (False, True)
                  if not (u and t) and (g == u)]      # the callers' combinations (:182-192, :205-216)


def _dense_reference(ei, ew, n, gcn_norm, set_diag, remove_diag, symmetrize, transpose):
    """float64 dense restatement of lib/sgp_preprocessing.py:67-105 (+ :182-185, :205-207)."""
    ei = np.asarray(ei)
    if transpose:
        ei = ei[[1, 0]]
    w = np.ones(ei.shape[1]) if ew is None else np.asarray(ew, np.float64)
    A = np.zeros((n, n))
    if symmetrize:                                   # to_undirected: both directions, coalesced by add
        np.add.at(A, (ei[0], ei[1]), w)
        np.add.at(A, (ei[1], ei[0]), w)
        if ew is None:                               # no attribute to add: coalesce only de-duplicates
            A = (A > 0).astype(np.float64)
        A = A.T                                      # then col, row = edge_index
    else:
        np.add.at(A, (ei[1], ei[0]), w)              # row = edge_index[1], col = edge_index[0]; duplicates add
    if set_diag:
        np.fill_diagonal(A, 1.0)
    elif remove_diag:
        np.fill_diagonal(A, 0.0)
    deg = A.sum(1)
    with np.errstate(divide="ignore"):
        if gcn_norm:
            d = deg ** -0.5
            d[np.isinf(d)] = 0
            return d[:, None] * A * d[None, :]
        d = deg ** -1.0
        d[np.isinf(d)] = 0
        return d[:, None] * A


@pytest.mark.parametrize("flags", ALL_FLAG_CASES, ids=lambda f: "-".join(k for k, v in f.items() if v) or "plain")
@pytest.mark.parametrize("weighted", [True, False])
def test_operator_all_flags_vs_dense_scipy_and_torch_sparse(flags, weighted):
    import scipy.sparse as sp
    n = 37
    g = np.random.default_rng(11)
    ei = g.integers(0, n, size=(2, 260)).astype(np.int64)
    ei = ei[:, ei[1] != 5]                                        # an empty row (degree 0 -> zero row)
    ei[:, :6] = np.array([[3, 3, 3, 7, 7, 9], [9, 9, 9, 7, 7, 3]])   # duplicates, a stored diagonal, a 2-cycle
    ew = g.uniform(0.1, 1.0, size=ei.shape[1]).astype(np.float32) if weighted else None
    e2, w2 = ei, ew
    if flags["transpose"]:
        e2 = e2[[1, 0]]
    if flags["symmetrize"]:
        e2, w2 = O.undirected_edges(e2, w2, n)
    rowptr, col, val = O.build_operator(e2, w2, n, gcn_norm=flags["gcn_norm"], set_diag=flags["set_diag"],
                                        remove_diag=flags["remove_diag"])
    want = _dense_reference(ei, ew, n, **flags)
    np.testing.assert_allclose(O.csr_to_dense(rowptr, col, val, n), want, rtol=2e-6, atol=1e-7)
    assert np.all(np.diff(col.astype(np.int64) + n * np.repeat(np.arange(n), np.diff(rowptr))) >= 0)   # (row, col) order
    x = g.standard_normal((n, 6)).astype(np.float32)
    S = sp.csr_matrix((val, col, rowptr), shape=(n, n))
    St = torch.sparse_csr_tensor(torch.from_numpy(rowptr), torch.from_numpy(col), torch.from_numpy(val), (n, n),
                                 check_invariants=False)      # duplicates are kept (torch_sparse semantics)
    y_np = want @ x.astype(np.float64)
    for got in (S @ x, (St @ torch.from_numpy(x)).numpy(), O.spmm_loops(rowptr, col, val, x[None])[0],
                O.spmm(rowptr, col, val, x[None], impl="c")[0]):
        np.testing.assert_allclose(got, y_np, rtol=1e-5, atol=1e-6)


# ---- SURVEY.md 8(f) restatements: hand-computed anchors --------------------------------------
def test_spatial_support_quirks():
    """[S, S^2, S^2] (never S^3), bidirectional = the row-normalised forward adjacency again, the
    dense 1/N matrix last (lib/sgp_preprocessing.py:143-158)."""
    A, ei, ew = path3()
    sup = O.spatial_support_dense(ei, ew, 3, k=3, bidirectional=True, global_attr=True)
    assert len(sup) == 7
    S = sup[0]
    np.testing.assert_allclose(sup[1], S @ S)
    np.testing.assert_allclose(sup[2], S @ S)                     # not S^3
    for a, b in zip(sup[:3], sup[3:6]):
        np.testing.assert_allclose(a, b)                          # the un-transposed recursion
    np.testing.assert_allclose(sup[6], np.full((3, 3), 1 / 3))
    und = O.spatial_support_dense(ei, ew, 3, k=1, undirected=True, bidirectional=True)
    assert und[0].shape == (3, 3) and not np.allclose(und[0], und[1])   # gcn norm vs row norm of A + A^T
    np.testing.assert_allclose(und[1], (A + A.T) / (A + A.T).sum(1, keepdims=True))
    np.testing.assert_allclose(S, A / A.sum(1, keepdims=True))


def test_iid_sample_and_grouped_conv_shapes():
    g = np.random.default_rng(0)
    x, y = g.standard_normal((20, 5, 6)), g.standard_normal((20, 5, 2))
    step, node = np.array([0, 3, 16]), np.array([4, 0, 2])
    xs, ys = O.iid_sample(x, y, step, node, horizon=3)
    assert xs.shape == (3, 1, 1, 6) and ys.shape == (3, 3, 1, 2)
    np.testing.assert_array_equal(xs[1, 0, 0], x[3, 0])
    np.testing.assert_array_equal(ys[2, :, 0], y[17:20, 2])
    w, b = g.standard_normal((6, 2, 1)), g.standard_normal(6)
    out = O.grouped_conv1x1(x[:2], w, b, groups=3)
    want = np.stack([x[:2, :, 2 * gi:2 * gi + 2] @ w[2 * gi:2 * gi + 2, :, 0].T + b[2 * gi:2 * gi + 2]
                     for gi in range(3)], -2).reshape(2, 5, 6)
    np.testing.assert_allclose(out, want, rtol=1e-12)


def test_gesn_operator_and_step_closed_form():
    ei = np.array([[0, 1, 1], [1, 0, 1]])                 # 0 -> 1, 1 -> 0, a stored loop on node 1; node 2 isolated
    S = O.gesn_operator_dense(ei, np.array([2.0, 3.0, 5.0]), 3)
    # loops appended for nodes 0..1 only (max index + 1); row 1 = {from 0: 2, loop: 5 + 1}; row 2 empty
    np.testing.assert_allclose(S, [[1 / 4, 3 / 4, 0], [2 / 8, 6 / 8, 0], [0, 0, 0]])
    layer = dict(w_ih=torch.tensor([[1.0]]), w_hh=torch.tensor([[0.5]]), b_ih=torch.tensor([0.0]), alpha=1.0)
    x = np.ones((2, 3, 1))
    out = O.graph_esn_states(x, [layer], S, "identity")
    np.testing.assert_allclose(out[0, :, 0], [1, 1, 1])
    np.testing.assert_allclose(out[1, :, 0], 1 + 0.5 * (S @ np.ones(3)))
