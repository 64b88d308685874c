"""Shared test helpers: golden loading and small seeded inputs."""
import ast
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(p)[len("reservoir_"):-len(".npz")]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, "reservoir_*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, f"reservoir_{name}.npz"))
    kwargs = ast.literal_eval(str(z["kwargs"]))
    L = kwargs.get("num_layers", 1)
    layers = [dict(w_ih=torch.from_numpy(z[f"w_ih_{i}"]), w_hh=torch.from_numpy(z[f"w_hh_{i}"]),
                   b_ih=torch.from_numpy(z[f"b_ih_{i}"]), alpha=float(z[f"alpha_{i}"]))
              for i in range(L)]
    return dict(seed=int(z["seed"]), x=z["x"], y=z["y"], kwargs=kwargs, layers=layers)


def random_graph(n, e, seed, weighted=True, self_loops=True):
    g = np.random.default_rng(seed)
    ei = g.integers(0, n, size=(2, e)).astype(np.int64)
    if not self_loops:
        ei = ei[:, ei[0] != ei[1]]
    ew = g.uniform(0.1, 1.0, size=ei.shape[1]).astype(np.float32) if weighted else None
    return ei, ew
