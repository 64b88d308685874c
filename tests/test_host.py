"""CPU-only checks of the host layer: ABI surface, weight generation parity with the reference's
golden weights, API mirror (signatures, argparse hooks, errors), row grouping (host code)."""
import argparse
import ctypes
import inspect
import os
import re

import numpy as np
import pytest
import torch

import sgp_b200
from sgp_b200 import _lib, ops
from sgp_b200.synthetic import sensor_knn, sensor_signal, sensor_thresh
from tests.helpers import golden_names, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "sgp_b200.h")).read()
    declared = set(re.findall(r"^(?:int|size_t|int64_t|const char\*)\s+(sgp_[a-z_0-9]+)\(", header, re.M))
    assert declared, "no declarations found in the header"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/sgp_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.sgp_version() >= 100


def test_no_cpu_fallback_product_raises_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    enc = sgp_b200.SGPTemporalEncoder(input_size=1, reservoir_size=8)
    with pytest.raises(_lib.SgpError):
        enc(torch.zeros(3, 2, 1))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "sgp_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("no CPU", ""), f"sgp_b200/{fn} mentions the oracle"


@pytest.mark.parametrize("name", golden_names())
def test_weight_init_bit_exact_vs_reference(name):
    g = load_golden(name)
    kw = dict(g["kwargs"])
    torch.manual_seed(g["seed"])
    res = sgp_b200.Reservoir(**kw)
    assert len(res.reservoir_layers) == len(g["layers"])
    for layer, ref in zip(res.reservoir_layers, g["layers"]):
        assert torch.equal(layer.w_ih.data, ref["w_ih"])
        assert torch.equal(layer.w_hh.data, ref["w_hh"])
        assert torch.equal(layer.b_ih.data, ref["b_ih"])
        assert float(layer.alpha) == ref["alpha"]
        assert not layer.w_hh.requires_grad


def test_reference_quirks_kept():
    with pytest.raises(ValueError):                       # tests/golden/reference_facts.txt
        sgp_b200.Reservoir(1, 4, activation="identity")
    with pytest.raises(AssertionError):
        sgp_b200.Reservoir(1, 4, activation="gelu")
    layer = sgp_b200.ReservoirLayer(2, 4, 0.9, 0.9, bias=False)
    assert layer.b_ih is not None                         # reservoir.py:47 `bias is not None`


def test_constructor_signatures_match_reference():
    want = {
        sgp_b200.SGPEncoder: ['input_size', 'reservoir_size', 'reservoir_layers', 'leaking_rate',
                              'spectral_radius', 'density', 'input_scaling', 'receptive_field',
                              'bidirectional', 'alpha_decay', 'global_attr', 'add_self_loops',
                              'undirected', 'reservoir_activation'],
        sgp_b200.SGPSpatialEncoder: ['receptive_field', 'bidirectional', 'undirected', 'global_attr',
                                     'add_self_loops'],
        sgp_b200.SGPTemporalEncoder: ['input_size', 'reservoir_size', 'reservoir_layers',
                                      'leaking_rate', 'spectral_radius', 'density', 'input_scaling',
                                      'alpha_decay', 'reservoir_activation'],
        sgp_b200.Reservoir: ['input_size', 'hidden_size', 'input_scaling', 'num_layers',
                             'leaking_rate', 'spectral_radius', 'density', 'activation', 'bias',
                             'alpha_decay'],
    }
    for cls, names in want.items():
        assert inspect.getfullargspec(cls.__init__).args[1:] == names, cls
    sig = inspect.signature(sgp_b200.sgp_spatial_embedding)
    assert list(sig.parameters) == ['x', 'num_nodes', 'edge_index', 'edge_weight', 'k', 'undirected',
                                    'add_self_loops', 'remove_self_loops', 'bidirectional',
                                    'one_hot_encoding', 'dropout_rate']
    assert sig.parameters['k'].default == 2
    sig = inspect.signature(sgp_b200.preprocess_adj)
    assert list(sig.parameters) == ['edge_index', 'edge_weight', 'num_nodes', 'gcn_norm', 'set_diag',
                                    'remove_diag']
    assert sig.parameters['set_diag'].default is True
    sig = inspect.signature(sgp_b200.encode_dataset)
    assert list(sig.parameters) == ['dataset', 'encoder_class', 'encoder_kwargs', 'encode_exogenous',
                                    'keep_raw', 'save_path']


@pytest.mark.parametrize("cls", [sgp_b200.SGPEncoder, sgp_b200.SGPSpatialEncoder,
                                 sgp_b200.SGPTemporalEncoder])
def test_argparse_hook(cls):
    p = cls.add_model_specific_args(argparse.ArgumentParser())
    ns = p.parse_args(['--receptive-field', '3', '--bidirectional', 'true'])
    assert ns.receptive_field == 3 and ns.bidirectional is True and ns.global_attr is False


def test_output_size_and_block_count():
    torch.manual_seed(0)
    enc = sgp_b200.SGPEncoder(3, 64, 2, 0.9, 0.9, 0.7, 1.0, 4, True, True, True)
    assert enc.output_size == 1280          # config/traffic/sgp_la.yaml: (1+4*2+1)*2*64
    assert enc.sgp_encoder.num_blocks() == 10


def test_bad_edge_index_type_raises_runtime_error():
    with pytest.raises(RuntimeError, match="Edge index must be"):
        sgp_b200.preprocess_adj([[0], [1]], None, 2)


def test_encode_dataset_nonbool_exogenous_is_nameerror():
    class D:
        exogenous = {}
    with pytest.raises(NameError):
        sgp_b200.encode_dataset(D(), sgp_b200.SGPTemporalEncoder, {}, encode_exogenous=["u"])


# ---- host-side row grouping (pure CPU code inside the .so) ---------------------------------
def _csr_from_edges(ei, w, n):
    import scipy.sparse as sp
    A = sp.csr_matrix((w, (ei[1], ei[0])), shape=(n, n))
    A.sort_indices()
    return A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float32)


@pytest.mark.parametrize("R", [4, 8, 16])
def test_group_rows_is_a_partition_and_compresses_knn(R):
    n, k = 3000, 24
    ei, w = sensor_knn(n, k, seed=3)
    rowptr, col, val = _csr_from_edges(ei, w, n)
    g = ops.group_rows_host(rowptr, col, val, n, R)
    assert g.shape == ((n + R - 1) // R, R)
    rows = g[g >= 0]
    assert rows.size == n and np.unique(rows).size == n
    grp = np.empty(n, np.int64)
    grp[g.reshape(-1)[g.reshape(-1) >= 0]] = (np.arange(g.size) // R)[g.reshape(-1) >= 0]
    e_rows = np.repeat(np.arange(n), np.diff(rowptr))
    union = np.unique(grp[e_rows] * n + col).size
    assert union < 0.75 * len(col)          # neighbours are shared inside groups


def test_group_rows_edge_cases():
    # empty graph, graph with empty rows, N not a multiple of R
    g = ops.group_rows_host(np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float32), 0, 4)
    assert g.shape == (0, 4)
    rowptr = np.array([0, 0, 2, 2, 3, 3], np.int32)
    col = np.array([0, 3, 1], np.int32)
    val = np.ones(3, np.float32)
    g = ops.group_rows_host(rowptr, col, val, 5, 4)
    assert g.shape == (2, 4) and sorted(g[g >= 0].tolist()) == [0, 1, 2, 3, 4]
    assert (g == -1).sum() == 3


def test_synthetic_generators_shapes():
    ei, w = sensor_knn(500, 10, seed=0)
    assert ei.shape == (2, 5000) and w.shape == (5000,) and ei.dtype == np.int64
    assert np.all(np.bincount(ei[1], minlength=500) == 10)       # exactly k entries per ROW
    assert not np.any(ei[0] == ei[1])
    ei, w = sensor_thresh(207, 1515, seed=0)
    assert abs(ei.shape[1] - 1515) < 60 and w.min() > 0.1
    x = sensor_signal(50, 7)
    assert x.shape == (50, 7, 3) and x.dtype == np.float32
    np.testing.assert_allclose(x[:, 0, 1], x[:, 3, 1])             # exogenous broadcast over nodes


def test_group_rows_has_no_scattered_leftover_groups():
    """Every 64-row group of a kNN graph must be a compact blob: the union of its rows' columns
    stays near the median (a deferred-leftovers grouping produced 10x outliers that cost 20% of
    the hop)."""
    from sgp_b200 import ops
    from sgp_b200.synthetic import sensor_knn
    n, k, R = 6000, 40, 64
    ei, ew = sensor_knn(n, k, seed=3)
    row, col = ei[1].astype(np.int64), ei[0].astype(np.int64)
    order = np.argsort(row * n + col, kind="stable")
    row, col, val = row[order], col[order], ew[order].astype(np.float32)
    rowptr = np.zeros(n + 1, np.int64)
    np.add.at(rowptr, row + 1, 1)
    rowptr = np.cumsum(rowptr)
    g = ops.group_rows_host(rowptr.astype(np.int32), col.astype(np.int32), val, n, R)
    assert sorted(g[g >= 0].tolist()) == list(range(n))
    assert (g[:-1] >= 0).all()                     # only the last group may be short
    unions = np.array([len(np.unique(np.concatenate([col[rowptr[r]:rowptr[r + 1]] for r in gr[gr >= 0]])))
                       for gr in g])
    assert unions.max() <= 2.5 * np.median(unions), (unions.max(), np.median(unions))


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with
    the contract's keys, the oracle port as cpu_baseline, no GPU work."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference",
                          "--workload", "c1_metr_la", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert key in line, key
    assert line["impl"] == "reference" and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["value"] > 0 and "workload" in line["config"]


def test_abi_version_matches_header_and_is_checked():
    lib = _lib.load()
    assert lib.sgp_version() == _lib.header_abi_version() >= 200


def test_chunk_rounding_and_sparse_adj_container():
    from sgp_b200.preprocessing import SparseAdj, _is_sparse_adj, _sparse_to_edges, round_chunk_steps
    assert [round_chunk_steps(s, 1000) for s in (1, 3, 4, 5, 16, 19, 1000, 5000)] == [1, 3, 4, 4, 16, 16, 1000, 1000]
    adj = SparseAdj(row=torch.tensor([1, 2, 2]), col=torch.tensor([0, 0, 1]), value=torch.tensor([1., 2., 3.]))
    assert adj.sparse_sizes() == (3, 3) and _is_sparse_adj(adj) and not _is_sparse_adj(torch.zeros(2, 3))
    ei, w, n = _sparse_to_edges(adj)
    assert n == 3 and ei.tolist() == [[0, 0, 1], [1, 2, 2]] and w.tolist() == [1., 2., 3.]   # [0] = col (source)
    assert _sparse_to_edges(adj.t())[0].tolist() == [[1, 2, 2], [0, 0, 1]]


def test_new_surface_is_exported_and_has_no_cpu_path():
    for name in ("IIDSampler", "sgp_spatial_support", "sgp_collate_features", "GroupedPointwiseConv",
                 "GESNEncoder", "GraphESN", "GESNLayer", "SparseAdj", "OperatorChain", "MeanOperator"):
        assert hasattr(sgp_b200, name), name
    if not torch.cuda.is_available():
        with pytest.raises(_lib.SgpError):
            sgp_b200.IIDSampler(torch.zeros(10, 3, 4), torch.zeros(10, 3, 1), horizon=2)
    # GESNEncoder mirrors the reference constructor (lib/nn/encoders/dyn_gesn_encoder.py:11-21)
    assert list(inspect.signature(sgp_b200.GESNEncoder.__init__).parameters)[1:] == [
        "input_size", "reservoir_size", "reservoir_layers", "leaking_rate", "spectral_radius", "density",
        "input_scaling", "alpha_decay", "reservoir_activation"]
    torch.manual_seed(0)
    enc = sgp_b200.GESNEncoder(2, 8, 2, 0.9, 0.9, 0.7, 1.0, True)
    assert [float(c.alpha) for c in enc.reservoir.rnn_cells] == [0.9, float(np.clip(0.9 - 0.1, 0.1, 1.))]
    p = argparse.ArgumentParser()
    sgp_b200.GESNEncoder.add_model_specific_args(p)
    assert p.parse_args(["--reservoir-size", "64", "--alpha-decay"]).alpha_decay is True


def test_small_reservoirs_choose_the_fused_multi_layer_plan():
    torch.manual_seed(0)
    assert sgp_b200.Reservoir(3, 64, num_layers=2).multi_layer_ok()            # sgp_la.yaml
    assert sgp_b200.Reservoir(3, 16, num_layers=8).multi_layer_ok()            # sgp_pv.yaml
    assert not sgp_b200.Reservoir(3, 128, num_layers=1).multi_layer_ok()       # tensor-core / tiled kernels
    assert not sgp_b200.Reservoir(3, 64, num_layers=8).multi_layer_ok()        # weights exceed shared memory


def test_to_host_passes_cpu_tensors_through():
    """ops.to_host stages CUDA tensors through pinned memory; CPU tensors come back as numpy views."""
    from sgp_b200 import ops
    a, b = torch.arange(5, dtype=torch.int32), torch.linspace(0, 1, 4)
    ha, hb = ops.to_host(a, b)
    assert isinstance(ha, np.ndarray) and ha.dtype == np.int32 and ha.tolist() == [0, 1, 2, 3, 4]
    assert hb.dtype == np.float32 and np.shares_memory(hb, b.numpy())
