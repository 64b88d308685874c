"""Generate tests/golden/reservoir_*.npz by executing the UNMODIFIED reference file
``/root/reference/lib/nn/reservoir/reservoir.py`` (commit ae1abd55).

Runs only where /root/reference exists (the build container); the .npz files are committed so
the tests never need the reference at run time.

``import lib.nn.reservoir`` cannot be used: the package __init__ pulls graph_reservoir.py ->
torch_geometric (not installed).  Instead the single file is loaded with
importlib.util.spec_from_file_location after registering two tiny stub modules for the only
names it imports from the rest of the reference:
  * tsl.nn.utils.get_functional_activation   (semantics of tsl/nn/utils/utils.py:34-44)
  * lib.utils.self_normalizing_activation    (semantics of lib/utils.py:50-51)

    python tests/golden/make_golden.py [case ...]
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference/lib/nn/reservoir/reservoir.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_reservoir():
    def get_functional_activation(name=None):
        if name is None or name.lower() == "linear":
            return lambda x: x
        name = name.lower()
        if name in ("tanh", "sigmoid"):
            return getattr(torch, name)
        if name == "identity":
            # tsl maps 'identity' -> F.identity?  tsl/nn/utils/utils.py has no such entry; the
            # reference's own assert allows it but get_functional_activation would raise.
            raise ValueError("Activation 'identity' not valid.")
        return getattr(F, name)

    for name in ("tsl", "tsl.nn", "tsl.nn.utils", "lib", "lib.utils"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["tsl.nn.utils"].get_functional_activation = get_functional_activation
    sys.modules["lib.utils"].self_normalizing_activation = \
        lambda x, r=1.0: r * F.normalize(x, p=2, dim=-1)
    spec = importlib.util.spec_from_file_location("_ref_reservoir", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


CASES = {
    # name: (seed, T, N, Fin, kwargs of reference Reservoir)
    "tanh_l1": (11, 24, 9, 3, dict(hidden_size=16, num_layers=1, leaking_rate=0.9,
                                    spectral_radius=0.9, density=0.7, activation="tanh")),
    "tanh_l2_decay": (12, 20, 7, 3, dict(hidden_size=32, num_layers=2, leaking_rate=0.9,
                                          spectral_radius=0.9, density=0.7, activation="tanh",
                                          alpha_decay=True, input_scaling=1.5)),
    "tanh_l3_dense": (13, 12, 5, 1, dict(hidden_size=8, num_layers=3, leaking_rate=0.8,
                                          spectral_radius=0.7, density=1.0, activation="tanh",
                                          alpha_decay=True)),
    "relu_l1": (14, 16, 6, 2, dict(hidden_size=16, num_layers=1, leaking_rate=0.7,
                                    spectral_radius=0.8, density=0.8, activation="relu")),
    "selfnorm_l2": (15, 16, 6, 2, dict(hidden_size=16, num_layers=2, leaking_rate=0.9,
                                        spectral_radius=0.9, density=0.9,
                                        activation="self_norm")),
    "tanh_h64_la": (16, 48, 23, 3, dict(hidden_size=64, num_layers=2, leaking_rate=0.9,
                                         spectral_radius=0.9, density=0.7, activation="tanh",
                                         alpha_decay=True)),
    # the shapes the tensor-core scan covers (H = 128 / 256, one layer, more than one 128-node tile)
    "tanh_h128_tc": (17, 30, 150, 3, dict(hidden_size=128, num_layers=1, leaking_rate=0.9,
                                           spectral_radius=0.9, density=0.7, activation="tanh")),
    "tanh_h256_tc": (18, 24, 70, 1, dict(hidden_size=256, num_layers=1, leaking_rate=0.9,
                                          spectral_radius=0.9, density=0.7, activation="tanh")),
}


def main():
    ref = load_reference_reservoir()
    only = set(sys.argv[1:])           # optional: regenerate just the named cases
    for name, (seed, T, N, Fin, kw) in CASES.items():
        if only and name not in only:
            continue
        torch.manual_seed(seed)
        res = ref.Reservoir(input_size=Fin, **kw)
        g = torch.Generator().manual_seed(seed + 1000)
        x = torch.randn(1, T, N, Fin, generator=g)
        with torch.no_grad():
            y = res(x)[0]
        blob = dict(seed=seed, x=x[0].numpy(), y=y.numpy(),
                    kwargs=np.array(repr(dict(input_size=Fin, **kw))))
        for i, layer in enumerate(res.reservoir_layers):
            blob[f"w_ih_{i}"] = layer.w_ih.numpy()
            blob[f"w_hh_{i}"] = layer.w_hh.numpy()
            blob[f"b_ih_{i}"] = layer.b_ih.numpy()
            blob[f"alpha_{i}"] = np.float64(layer.alpha)
        path = os.path.join(HERE, f"reservoir_{name}.npz")
        np.savez_compressed(path, **blob)
        print(f"{name}: y{tuple(y.shape)} -> {path} ({os.path.getsize(path)} B)")
    # identity activation: the reference asserts it is allowed (reservoir.py:37) but its
    # get_functional_activation raises on it (tsl/nn/utils/utils.py:34-44) -> record that fact.
    try:
        ref.Reservoir(input_size=1, hidden_size=4, activation="identity")
        ident = "constructs"
    except Exception as e:  # noqa: BLE001
        ident = f"{type(e).__name__}: {e}"
    with open(os.path.join(HERE, "reference_facts.txt"), "w") as f:
        f.write(f"Reservoir(activation='identity') in the reference: {ident}\n")
    print("identity:", ident)


if __name__ == "__main__":
    main()
