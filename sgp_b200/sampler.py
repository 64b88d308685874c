"""IID (t, n) sampling from DEVICE-RESIDENT embeddings — the step right after the encoder.

Reference: ``IIDDataset.sample`` (lib/datasets/iid_dataset.py:57-99) driven by ``IIDLoader``
(lib/dataloader/iid_dataloader.py:25-46).  There every batch is cut out of a host tensor by
DataLoader workers (``tens[(step_index, None, None, node_index)]``) and copied to the GPU; here
the encoder's output stays on the device (one [T, N, D] tensor, or this rank's rows of it) and a
batch is ONE gather kernel (``sgp_gather_tn``) per tensor: no host round trip, constant cost per
batch whatever T and N are.

Same index arithmetic and, by default, the same host RNG calls in the same order
(``torch.randint(0, n_steps - horizon, (B,))`` then ``torch.randint(0, n_nodes, (B,))``), so a
seeded run draws exactly the reference's samples; ``device_rng=True`` draws on the GPU instead
(no H2D copy of the indices, different stream of random numbers).
"""
from __future__ import annotations

from typing import Dict, Iterator, Optional

import torch
from torch import Tensor

from . import ops
from ._lib import SgpError


class IIDSampler:
    """Random (step, node) mini-batches of encoder output ``x`` [T, N, D] and targets ``y``
    [T, N, C] (both CUDA, float32).  ``u`` [T, Fu]: optional time-only exogenous (pattern 't f').

    ``sample(B)`` returns ``dict(x=[B, 1, 1, D], y=[B, h, 1, C], node_index=[B, 1][, u=[B, 1, Fu]])``
    with ``h = len(range(delay + 1, horizon + 1, horizon_lag))`` — the shapes of the reference's
    ``Data.input.x`` / ``Data.target.y`` / ``Data.input.node_index``."""

    def __init__(self, x: Tensor, y: Tensor, horizon: int, delay: int = 0, horizon_lag: int = 1,
                 u: Optional[Tensor] = None, batch_size: int = 1024, num_batches: int = 1000,
                 device_rng: bool = False):
        if not (x.is_cuda and y.is_cuda):
            raise SgpError("IIDSampler gathers from device-resident tensors (there is no CPU path)")
        if x.dim() != 3 or y.dim() != 3 or x.shape[:2] != y.shape[:2]:
            raise ValueError(f"x {tuple(x.shape)} / y {tuple(y.shape)}: expected [T, N, D] and [T, N, C]")
        self.x, self.y, self.u = x, y, u
        self.n_steps, self.n_nodes = int(x.shape[0]), int(x.shape[1])
        self.horizon, self.delay, self.horizon_lag = int(horizon), int(delay), int(horizon_lag)
        self.offsets = list(range(self.delay + 1, self.horizon + 1, self.horizon_lag))
        if self.n_steps - self.horizon <= 0:
            raise ValueError("series shorter than the horizon")
        self.batch_size, self.num_batches, self.device_rng = int(batch_size), int(num_batches), device_rng
        self._off_dev = torch.tensor(self.offsets, dtype=torch.int64, device=x.device)

    def draw_indices(self, B: int):
        """(step_index [B], node_index [B]) int64 on the device."""
        dev = self.x.device
        if self.device_rng:
            return (torch.randint(0, self.n_steps - self.horizon, (B,), device=dev),
                    torch.randint(0, self.n_nodes, (B,), device=dev))
        step = torch.randint(0, self.n_steps - self.horizon, (B,))      # iid_dataset.py:58
        node = torch.randint(0, self.n_nodes, (B,))                      # iid_dataset.py:59
        return step.to(dev, non_blocking=True), node.to(dev, non_blocking=True)

    def gather(self, step_index: Tensor, node_index: Tensor) -> Dict[str, Tensor]:
        dev = self.x.device
        B, h = int(step_index.numel()), len(self.offsets)
        D, C = int(self.x.shape[-1]), int(self.y.shape[-1])
        out_x = torch.empty(B, D, device=dev)
        ops.gather_tn(self.x, step_index, node_index, out_x)
        # hor_index = stack([step + i for i in range(delay + 1, horizon + 1, lag)], 1)   (:80-82)
        hor = (step_index[:, None] + self._off_dev[None, :]).reshape(-1).contiguous()
        nodes = node_index[:, None].expand(B, h).reshape(-1).contiguous()
        out_y = torch.empty(B * h, C, device=dev)
        ops.gather_tn(self.y, hor, nodes, out_y)
        batch = dict(x=out_x.view(B, 1, 1, D), y=out_y.view(B, h, 1, C), node_index=node_index[:, None])
        if self.u is not None:
            batch["u"] = self.u.index_select(0, step_index)[:, None]     # tens[(step_index, None)]  (:69-70)
        return batch

    def sample(self, N: Optional[int] = None) -> Dict[str, Tensor]:
        step, node = self.draw_indices(int(N or self.batch_size))
        return self.gather(step, node)

    def __len__(self) -> int:
        return self.num_batches

    def __iter__(self) -> Iterator[Dict[str, Tensor]]:
        for _ in range(self.num_batches):
            yield self.sample(self.batch_size)
