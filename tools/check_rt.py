"""Determinism / accuracy soak of the tensor-core scan (debug aid): repeated launches must be
bit-identical; the first is compared with the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import sgp_oracle as O
from sgp_b200 import ops
DEV = "cuda:0"
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for (H, N, Fin, T) in [(256, 300, 1, 50), (256, 129, 3, 50), (128, 700, 3, 50), (256, 20000, 1, 12)]:
    torch.manual_seed(H + N + Fin)
    layers = O.draw_reservoir(Fin, H, 1, 0.9, 0.9, 0.7)
    l = layers[0]
    x = torch.as_tensor(np.random.default_rng(N).standard_normal((T, N, Fin)).astype(np.float32), device=DEV)
    wimg = ops.reservoir_tc_pack(l["w_hh"].to(DEV))
    w_ih, b = l["w_ih"].to(DEV).contiguous(), l["b_ih"].to(DEV)
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    first = None; bad = 0
    for rep in range(reps):
        out = torch.full((T, N, H), float("nan"), device=DEV)
        state = torch.zeros(N, H, device=DEV)
        ops.reservoir_scan_tc(x, wimg, w_ih, b, l["alpha"], "tanh", state, out, err)
        if first is None:
            first = out.clone()
            if N <= 1000:
                ref = O.reservoir_states(x.cpu().numpy(), layers, "tanh").numpy()
                ref64 = O.reservoir_states(x.cpu().numpy(), layers, "tanh", dtype=torch.float64).numpy()
                wp = ops.reservoir_pack(l["w_ih"].to(DEV), l["w_hh"].to(DEV))
                o2 = torch.empty(T, N, H, device=DEV); s2 = torch.zeros(N, H, device=DEV)
                ops.reservoir_scan(x, wp, b, l["alpha"], "tanh", s2, o2)
                y = first.cpu().numpy(); y2 = o2.cpu().numpy()
                e = np.abs(y - ref64)
                print(f"H={H} N={N} Fin={Fin}: vs f64 oracle: tc {e.max():.3e}  cuda-core {np.abs(y2 - ref64).max():.3e}  "
                      f"f32 oracle {np.abs(ref - ref64).max():.3e}")
                print("   tc err by t     :", " ".join(f"{v:.0e}" for v in e.max(axis=(1, 2))[:16]))
                print("   tc err by chunk :", " ".join(f"{e[:, :, c * 32:(c + 1) * 32].max():.0e}" for c in range(H // 32)))
                print("   tc err by rows/32:", " ".join(f"{e[:, r:r + 32].max():.0e}" for r in range(0, N, 32))[:200])
        else:
            diff = (out != first) | torch.isnan(out)
            if bool(diff.any()):
                bad += 1
                idx = diff.nonzero()
                d = (out - first).abs()
                print(f"  rep {rep}: {idx.shape[0]} elements differ, max {float(d[diff].max()):.3e}; "
                      f"t {sorted(set(idx[:, 0].tolist()))[:8]} rows {int(idx[:, 1].min())}..{int(idx[:, 1].max())} "
                      f"cols {int(idx[:, 2].min())}..{int(idx[:, 2].max())}")
                if bad > 5:
                    break
    print(f"H={H} N={N}: {bad} of {reps} launches differ from the first; err flag {int(err.item())}")
