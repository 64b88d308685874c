"""The encoder classes — the reference's plugin surface (lib/nn/encoders/*), backed by the
sm_100a kernels.

Same class names, ``__init__`` argument names (``tsl``'s ``filter_args`` routes CLI args to
constructors by parameter NAME), attributes (``.reservoir``, ``.sgp_encoder``),
``forward(x, edge_index, edge_weight)`` and ``add_model_specific_args`` as
``SGPEncoder`` (sgp_encoder.py:9-80), ``SGPSpatialEncoder`` (sgp_spatial_encoder.py:8-50) and
``SGPTemporalEncoder`` (sgp_temporal_encoder.py:8-60).

Execution model: the series is cut into time chunks; per chunk the input is copied to the GPU,
every reservoir layer is ONE fused scan launch writing its states straight into feature block 0
of the chunk's ``[Tc, N, D]`` output buffer, every hop is one SpMM launch writing the next
block of the same buffer, and the finished chunk goes to the caller's sink (a host tensor for the
reference-compatible ``forward``, or any callable for device-resident / streaming use).  The
reservoir state is carried across chunks, so the result does not depend on the chunking.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
from torch import nn, Tensor

from . import ops
from .preprocessing import (ShiftOperator, _chunk_steps, make_operators, panel_bound, propagate_into,
                            spatial_blocks)
from .reservoir import Reservoir, _cuda_device_for


def _str_to_bool(value):
    if isinstance(value, bool):
        return value
    if value.lower() in {'false', 'f', '0', 'no', 'n', 'off'}:
        return False
    if value.lower() in {'true', 't', '1', 'yes', 'y', 'on'}:
        return True
    raise ValueError(f'{value} is not a valid boolean value')


def _opt_list(parser, *args, **kwargs):
    """test_tube's ``opt_list`` when the parser has it (the reference's HyperOptArgumentParser),
    plain ``add_argument`` otherwise."""
    if hasattr(parser, "opt_list"):
        parser.opt_list(*args, **kwargs)
    else:
        kwargs.pop("tunable", None)
        kwargs.pop("options", None)
        parser.add_argument(*args, **kwargs)


def _add_reservoir_args(parser):
    _opt_list(parser, '--reservoir-size', type=int, default=32, tunable=True,
              options=[16, 32, 64, 128, 256])
    _opt_list(parser, '--reservoir-layers', type=int, default=1, tunable=True, options=[1, 2, 3])
    _opt_list(parser, '--spectral-radius', type=float, default=0.9, tunable=True,
              options=[0.7, 0.8, 0.9])
    _opt_list(parser, '--leaking-rate', type=float, default=0.9, tunable=True,
              options=[0.7, 0.8, 0.9])
    _opt_list(parser, '--density', type=float, default=0.7, tunable=True, options=[0.7, 0.8, 0.9])
    _opt_list(parser, '--input-scaling', type=float, default=1., tunable=True,
              options=[1., 1.5, 2.])
    _opt_list(parser, '--alpha-decay', type=_str_to_bool, nargs='?', const=True, default=False)
    parser.add_argument('--reservoir-activation', type=str, default='tanh')


def _add_spatial_args(parser):
    _opt_list(parser, '--receptive-field', type=int, default=1, tunable=True, options=[1, 2, 3])
    _opt_list(parser, '--bidirectional', type=_str_to_bool, nargs='?', const=True, default=False)
    _opt_list(parser, '--undirected', type=_str_to_bool, nargs='?', const=True, default=False)
    _opt_list(parser, '--add-self-loops', type=_str_to_bool, nargs='?', const=True, default=False)
    _opt_list(parser, '--global-attr', type=_str_to_bool, nargs='?', const=True, default=False)


class SGPSpatialEncoder(nn.Module):
    """K hops of the shift operator (+ optional reversed graph, + optional global-mean block),
    concatenated on the feature axis.  Reference: sgp_spatial_encoder.py:8-35."""

    def __init__(self, receptive_field, bidirectional, undirected, global_attr, add_self_loops=False):
        super().__init__()
        self.receptive_field = receptive_field
        self.bidirectional = bidirectional
        self.undirected = undirected
        self.add_self_loops = add_self_loops
        self.global_attr = global_attr
        self.rbu_mode = "auto"      # "auto" | "off" | "force4/8/16" (B200-side tuning knob)

    # ---- shape helpers ----
    def num_blocks(self) -> int:
        return spatial_blocks(self.receptive_field, self.bidirectional) + (1 if self.global_attr else 0)

    def build_operators(self, edge_index, edge_weight, num_nodes, device, F):
        return make_operators(edge_index, edge_weight, num_nodes, undirected=self.undirected,
                              add_self_loops=self.add_self_loops, remove_self_loops=False,
                              bidirectional=self.bidirectional, device=device, F=F,
                              rbu=self.rbu_mode)

    def encode_chunk(self, buf: Tensor, F: int, fwd: ShiftOperator, bwd: Optional[ShiftOperator],
                     sums: Optional[Tensor] = None, checksum: Optional[Tensor] = None,
                     bound: Optional[float] = None) -> None:
        """buf [Tc, N, num_blocks*F] (device) with block 0 filled; fills the other blocks.
        `checksum` (device float64 scalar) += the sum of every block written here."""
        k = self.receptive_field
        propagate_into(buf, F, k, fwd, bwd, checksum, bound)
        if self.global_attr:
            Tc, N, _ = buf.shape
            if sums is None:
                sums = torch.empty(Tc, F, device=buf.device)
            ops.node_sum(buf[..., :F], sums[:Tc])
            g = spatial_blocks(k, self.bidirectional)
            ops.node_mean_broadcast(sums[:Tc], N, buf[..., g * F:(g + 1) * F])
            if checksum is not None:
                ops.checksum_view(buf[..., g * F:(g + 1) * F], checksum)

    def forward(self, x, edge_index, edge_weight):
        """x [T, N, F] (or [N, F]) on CPU or GPU -> [T, N, num_blocks*F] on the same device."""
        squeeze = x.dim() == 2
        x3 = x[None] if squeeze else x
        T, N, F = x3.shape
        dev = _cuda_device_for(x3)
        fwd, bwd = self.build_operators(edge_index, edge_weight, N, dev, F)
        D = self.num_blocks() * F
        out = torch.empty(T, N, D, dtype=torch.float32, device=x.device)
        step = T if x.is_cuda else _chunk_steps(T, N * D * 4)
        for t0 in range(0, T, step):
            t1 = min(T, t0 + step)
            buf = out[t0:t1] if x.is_cuda else torch.empty(t1 - t0, N, D, device=dev)
            buf[..., :F] = x3[t0:t1].to(device=dev, dtype=torch.float32)
            self.encode_chunk(buf, F, fwd, bwd, bound=panel_bound(buf[..., :F]) if fwd.tc16 is not None else None)
            if not x.is_cuda:
                out[t0:t1] = buf.to(x.device)
        for op in (fwd, bwd):
            if op is not None:
                op.check()
        return out[0] if squeeze else out

    @staticmethod
    def add_model_specific_args(parser):
        _add_spatial_args(parser)
        return parser


class SGPEncoder(nn.Module):
    """Reservoir over every node's series, then the spatial encoder.
    Reference: sgp_encoder.py:9-51."""

    def __init__(self, input_size, reservoir_size, reservoir_layers, leaking_rate, spectral_radius,
                 density, input_scaling, receptive_field, bidirectional, alpha_decay, global_attr,
                 add_self_loops=False, undirected=False, reservoir_activation='tanh'):
        super().__init__()
        self.reservoir = Reservoir(input_size=input_size, hidden_size=reservoir_size,
                                   input_scaling=input_scaling, num_layers=reservoir_layers,
                                   leaking_rate=leaking_rate, spectral_radius=spectral_radius,
                                   density=density, activation=reservoir_activation,
                                   alpha_decay=alpha_decay)
        self.sgp_encoder = SGPSpatialEncoder(receptive_field=receptive_field,
                                             bidirectional=bidirectional, undirected=undirected,
                                             add_self_loops=add_self_loops, global_attr=global_attr)
        self.chunk_steps: Optional[int] = None      # None = sized from SGP_B200_CHUNK_BYTES

    @property
    def output_size(self) -> int:
        return self.sgp_encoder.num_blocks() * self.reservoir.num_layers * self.reservoir.hidden_size

    def encode_stream(self, x: Tensor, edge_index, edge_weight,
                      sink: Optional[Callable[[int, int, Tensor], None]], device=None,
                      operators=None, checksum: Optional[Tensor] = None) -> None:
        """Encode x [T, N, Fin] chunk by chunk; ``sink(t0, t1, chunk)`` receives each finished
        [t1-t0, N, D] device buffer (valid until the next-but-one call: two buffers alternate).
        ``checksum`` (device float64 scalar) += the sum of the whole [T, N, D] output, accumulated
        by the producing kernels themselves; with it ``sink`` may be None (a pure streaming run)."""
        T, N, Fin = x.shape
        dev = torch.device(device) if device is not None else _cuda_device_for(x)
        res, spat = self.reservoir, self.sgp_encoder
        L, H = res.num_layers, res.hidden_size
        F, D = L * H, self.output_size
        fwd, bwd = operators if operators is not None else \
            spat.build_operators(edge_index, edge_weight, N, dev, F)
        plan = res.device_plan(dev, N)
        step = self.chunk_steps or _chunk_steps(T, N * D * 4)
        state = torch.zeros(L, N, H, device=dev)
        bufs = [torch.empty(step, N, D, device=dev) for _ in range(2 if step < T else 1)]
        sums = torch.empty(step, F, device=dev) if spat.global_attr else None
        for i, t0 in enumerate(range(0, T, step)):
            t1 = min(T, t0 + step)
            buf = bufs[i % len(bufs)][: t1 - t0]
            xc = x[t0:t1].detach().to(device=dev, dtype=torch.float32, non_blocking=True)
            res.scan_chunk(plan, xc, state, buf, checksum)
            spat.encode_chunk(buf, F, fwd, bwd, sums, checksum, bound=res.state_bound())
            if sink is not None:
                sink(t0, t1, buf)
        for op in (fwd, bwd):
            if op is not None:
                op.check()
        res.check_plan(plan)

    def forward(self, x, edge_index, edge_weight):
        """x [T, N, Fin] on CPU or GPU -> [T, N, D] on the same device."""
        T, N, _ = x.shape
        out = torch.empty(T, N, self.output_size, dtype=torch.float32, device=x.device)

        def sink(t0, t1, chunk):
            out[t0:t1].copy_(chunk)

        self.encode_stream(x, edge_index, edge_weight, sink)
        return out

    @staticmethod
    def add_model_specific_args(parser):
        _opt_list(parser, '--reservoir-size', type=int, default=32, tunable=True,
                  options=[16, 32, 64, 128, 256])
        _opt_list(parser, '--reservoir-layers', type=int, default=1, tunable=True, options=[1, 2, 3])
        _opt_list(parser, '--receptive-field', type=int, default=1, tunable=True, options=[1, 2, 3])
        _opt_list(parser, '--spectral-radius', type=float, default=0.9, tunable=True,
                  options=[0.7, 0.8, 0.9])
        _opt_list(parser, '--leaking-rate', type=float, default=0.9, tunable=True,
                  options=[0.7, 0.8, 0.9])
        _opt_list(parser, '--density', type=float, default=0.7, tunable=True, options=[0.7, 0.8, 0.9])
        _opt_list(parser, '--input-scaling', type=float, default=1., tunable=True,
                  options=[1., 1.5, 2.])
        _opt_list(parser, '--bidirectional', type=_str_to_bool, nargs='?', const=True, default=False)
        _opt_list(parser, '--undirected', type=_str_to_bool, nargs='?', const=True, default=False)
        _opt_list(parser, '--add-self-loops', type=_str_to_bool, nargs='?', const=True, default=False)
        _opt_list(parser, '--alpha-decay', type=_str_to_bool, nargs='?', const=True, default=False)
        _opt_list(parser, '--global-attr', type=_str_to_bool, nargs='?', const=True, default=False)
        parser.add_argument('--reservoir-activation', type=str, default='tanh')
        return parser


class SGPTemporalEncoder(nn.Module):
    """Reservoir only (the "no spatial encoding" ablation); graph arguments are ignored.
    Reference: sgp_temporal_encoder.py:8-34."""

    def __init__(self, input_size, reservoir_size=32, reservoir_layers=1, leaking_rate=0.9,
                 spectral_radius=0.9, density=0.7, input_scaling=1., alpha_decay=False,
                 reservoir_activation='tanh'):
        super().__init__()
        self.reservoir = Reservoir(input_size=input_size, hidden_size=reservoir_size,
                                   input_scaling=input_scaling, num_layers=reservoir_layers,
                                   leaking_rate=leaking_rate, spectral_radius=spectral_radius,
                                   density=density, activation=reservoir_activation,
                                   alpha_decay=alpha_decay)

    def forward(self, x: Tensor, *args, **kwargs):
        T, N, _ = x.shape
        dev = _cuda_device_for(x)
        res = self.reservoir
        L, H = res.num_layers, res.hidden_size
        plan = res.device_plan(dev, N)
        out = torch.empty(T, N, L * H, dtype=torch.float32, device=x.device)
        state = torch.zeros(L, N, H, device=dev)
        step = T if x.is_cuda else _chunk_steps(T, N * L * H * 4)
        for t0 in range(0, T, step):
            t1 = min(T, t0 + step)
            buf = out[t0:t1] if x.is_cuda else torch.empty(t1 - t0, N, L * H, device=dev)
            res.scan_chunk(plan, x[t0:t1].detach().to(device=dev, dtype=torch.float32), state, buf)
            if not x.is_cuda:
                out[t0:t1] = buf.to(x.device)
        res.check_plan(plan)
        return out

    @staticmethod
    def add_model_specific_args(parser):
        _add_reservoir_args(parser)
        _add_spatial_args(parser)   # the reference declares them here too (unused by this class)
        return parser
