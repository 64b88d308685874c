"""Seeded synthetic sensor graphs and signals of the BASELINE.json shapes.

There are no dataset files and no network, so the benchmark and the large parity tests use
graphs built the way the reference builds its own:

* ``sensor_knn``    — PV-US / CER-En recipe (experiments/run_largescale_sgp.py:167-170 on top of
  tsl/ops/similarities.py:58-62, 104-122): Gaussian kernel of the pairwise distance, each ROW
  keeps its k nearest neighbours, emitted in the reference's edge convention
  ``edge_index = [col (source j); row (target i)]`` (tsl/ops/connectivity.py:130-143).
* ``sensor_thresh`` — METR-LA / PEMS-BAY recipe (experiments/run_traffic_sgp.py:161): Gaussian
  kernel thresholded at 0.1, no self loops; irregular row lengths, possibly empty rows.
* ``sensor_signal`` — standardised-traffic-like series: daily sinusoid with a per-node phase plus
  noise, and optionally the two day-of-time exogenous channels broadcast over nodes
  (tsl/datasets/prototypes/mixin.py:97-115).

Everything is numpy/scipy on the host (one-off input generation, not part of the timed path).
"""
from __future__ import annotations

import numpy as np


def sensor_positions(n: int, seed: int = 0) -> np.ndarray:
    return np.random.default_rng(seed).random((n, 2))


def sensor_knn(n: int, k: int, seed: int = 0, permute: bool = True, return_pos: bool = False):
    """k-NN Gaussian-kernel graph on n uniform points of the unit square.

    Exactly k stored entries per row (target node i <- its k nearest sources j != i), weight
    exp(-(d/theta)^2) with theta = std of the kept distances.  ``permute`` randomly relabels
    the nodes so that node ids carry no spatial locality (the honest default)."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    pos = rng.random((n, 2))
    if permute:
        pos = pos[rng.permutation(n)]
    k = min(k, n - 1)
    tree = cKDTree(pos)
    dist, idx = tree.query(pos, k=k + 1, workers=-1)
    dist, idx = dist[:, 1:], idx[:, 1:]          # drop self
    theta = dist.std()
    w = np.exp(-np.square(dist / theta)).astype(np.float32)
    row = np.repeat(np.arange(n, dtype=np.int64), k)
    col = idx.reshape(-1).astype(np.int64)
    edge_index = np.stack([col, row])            # [0] = source j (column), [1] = target i (row)
    if return_pos:
        return edge_index, w.reshape(-1), pos
    return edge_index, w.reshape(-1)


def sensor_thresh(n: int, target_edges: int, seed: int = 0, threshold: float = 0.1):
    """Thresholded Gaussian-kernel graph (dense pairwise distances: small n only).  The kernel
    bandwidth is bisected so that about ``target_edges`` off-diagonal weights exceed the
    threshold; rows have irregular lengths and some may be empty."""
    rng = np.random.default_rng(seed)
    pos = rng.random((n, 2))
    d = np.sqrt(((pos[:, None, :] - pos[None, :, :]) ** 2).sum(-1))
    off = ~np.eye(n, dtype=bool)
    lo, hi = 1e-4, 2.0
    for _ in range(60):
        theta = 0.5 * (lo + hi)
        cnt = int(((np.exp(-np.square(d / theta)) > threshold) & off).sum())
        if cnt > target_edges:
            hi = theta
        else:
            lo = theta
    A = np.exp(-np.square(d / theta))
    A[~off] = 0.0
    A[A <= threshold] = 0.0
    jj, ii = np.nonzero(A.T)                     # reference: nonzero of the transposed adjacency
    edge_index = np.stack([jj, ii]).astype(np.int64)
    return edge_index, A.T[jj, ii].astype(np.float32)


def sensor_signal(T: int, n: int, seed: int = 1, exogenous: bool = True, period: int = 288,
                  t0: int = 0) -> np.ndarray:
    """[T, n, Fin] float32: channel 0 = sin(2 pi t / period + phi_n) + 0.3 N(0,1); with
    ``exogenous`` two more channels sin/cos(2 pi t / period) broadcast over the nodes (Fin = 3)."""
    rng = np.random.default_rng(seed)
    phase = rng.uniform(0, 2 * np.pi, size=n)
    t = (np.arange(t0, t0 + T, dtype=np.float64))[:, None]
    x = np.sin(2 * np.pi * t / period + phase[None, :]) + 0.3 * rng.standard_normal((T, n))
    if not exogenous:
        return x[..., None].astype(np.float32)
    day = 2 * np.pi * t / period
    u = np.concatenate([np.sin(day), np.cos(day)], -1)[:, None, :].repeat(n, 1)
    return np.concatenate([x[..., None], u], -1).astype(np.float32)


# name -> (N, graph kind, graph arg, T, H, K, Fin)   — BASELINE.json `configs`
CONFIGS = {
    "c1_metr_la": dict(N=207, graph="thresh", edges=1515, T=288, H=64, K=2, Fin=3),
    "c2_pems_bay": dict(N=325, graph="knn", k=8, T=52000, H=128, K=4, Fin=3),
    "c3_pv_us": dict(N=5016, graph="knn", k=100, T=10000, H=256, K=4, Fin=3),
    "c4_100k": dict(N=100_000, graph="knn", k=100, T=1000, H=256, K=4, Fin=1),
    "c5_1m": dict(N=1_000_000, graph="knn", k=32, T=256, H=128, K=2, Fin=1),
    # the reference's SHIPPED encoder configurations at their datasets' shapes (sgp_paper.pdf Table 3):
    # config/traffic/sgp_la.yaml (METR-LA 34272 x 207) and config/largescale_100nn/sgp_pv.yaml (PV-US 8868 x 5016)
    "la_yaml": dict(N=207, graph="thresh", edges=1515, T=34272, H=64, L=2, K=4, Fin=3, bidir=True, glob=True,
                    decay=True, leak=0.9, rho=0.9),
    "pv_yaml": dict(N=5016, graph="knn", k=100, T=8868, H=16, L=8, K=2, Fin=3, bidir=False, glob=True,
                    decay=True, leak=1.0, rho=0.99),
}


def make_graph(cfg: dict, seed: int = 0, permute: bool = True):
    if cfg["graph"] == "knn":
        return sensor_knn(cfg["N"], cfg["k"], seed=seed, permute=permute)
    return sensor_thresh(cfg["N"], cfg["edges"], seed=seed)
