"""Large-N check of the tensor-core path: BASELINE config C5 (N = 1 000 000, 32-NN, H = 128, K = 2)
on one GPU, a few time steps against the CPU oracle, then a timing of 32 steps.
    python tools/check_c5.py [N]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import sgp_b200
from oracle import sgp_oracle as O
from sgp_b200 import ops
from sgp_b200.synthetic import CONFIGS, make_graph, sensor_signal

cfg = dict(CONFIGS["c5_1m"])
if len(sys.argv) > 1:
    cfg["N"] = int(sys.argv[1])
N, H, K, Fin = cfg["N"], cfg["H"], cfg["K"], cfg["Fin"]
t0 = time.time()
ei, ew = make_graph(cfg, seed=0)
print(f"graph N={N} nnz={ei.shape[1]} in {time.time() - t0:.1f}s", flush=True)
torch.manual_seed(2)
enc = sgp_b200.SGPEncoder(input_size=Fin, reservoir_size=H, reservoir_layers=1, leaking_rate=0.9,
                          spectral_radius=0.9, density=0.7, input_scaling=1.0, receptive_field=K,
                          bidirectional=False, alpha_decay=False, global_attr=False)
T = 3
x = sensor_signal(T, N, seed=1, exogenous=Fin == 3)
dev = torch.device("cuda", 0)
t0 = time.time()
y = enc(torch.from_numpy(x).to(dev), torch.from_numpy(ei).to(dev), torch.from_numpy(ew).to(dev))
torch.cuda.synchronize()
print(f"encode T={T}: {time.time() - t0:.2f}s (includes operator build)", flush=True)
fwd, _ = enc.sgp_encoder.build_operators(torch.from_numpy(ei).to(dev), torch.from_numpy(ew).to(dev), N, dev, H)
print("hop format:", "tcgen05 fill %.3f" % fwd.tc.fill if fwd.tc is not None else "rbu" if fwd.rbu is not None else "csr",
      "| scan:", enc.reservoir.device_plan(dev, N)[0][0], flush=True)
layers = [dict(w_ih=l.w_ih.data, w_hh=l.w_hh.data, b_ih=l.b_ih.data, alpha=l.alpha) for l in enc.reservoir.reservoir_layers]
t0 = time.time()
ref = O.sgp_encoder(x, ei, ew, layers, "tanh", K, False, False, False, impl="c")
print(f"oracle: {time.time() - t0:.1f}s", flush=True)
ok, worst = O.blockwise_allclose(y.cpu().numpy(), ref, H)
print(f"C5 parity: worst |err|/tol = {worst:.3g} ({'OK' if ok else 'FAIL'})", flush=True)
# timing: 32 steps streamed in chunks, checksum sink
Tt = 32
xt = torch.from_numpy(sensor_signal(Tt, N, seed=1, exogenous=Fin == 3)).to(dev)
acc = torch.zeros(1, dtype=torch.float64, device=dev)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    enc.encode_stream(xt, torch.from_numpy(ei), torch.from_numpy(ew), lambda a, b, chunk: ops.checksum(chunk, acc))
    torch.cuda.synchronize(); dt = time.time() - t0
print(f"encode_stream T={Tt}: {dt:.2f}s wall incl. operator build = {N * Tt / dt / 1e6:.1f} M node-steps/s")
assert ok
