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
