"""First decoder layer over the encoder's feature blocks.

Reference: ``SGPModel.input_encoder`` (lib/nn/models/sgp_model.py:41-52):
``Rearrange('b n f -> b f n') -> nn.Conv1d(input_size, out_channels, kernel_size=1, groups=order)
-> Rearrange('b f n -> b n f')`` with ``out_channels = hidden_size - hidden_size % order`` and
``order = (1 + K (1|2) + global) * L`` — one group per (hop, layer) block of the encoder output, so
the layer is a block-diagonal linear map.  :class:`GroupedPointwiseConv` has the same parameters
(``weight [out_channels, input_size / groups, 1]``, ``bias [out_channels]``, the same default
initialisation as ``nn.Conv1d``) and takes ``[b, n, f]`` directly: the forward is ONE kernel
(``sgp_grouped_linear``) reading the sampled rows where they lie — no rearrange copies.  The layer is
trainable: gradients of weight / bias / input are formed with batched matmuls on the grouped views
(the decoder's training loop is outside the hot path).
"""
from __future__ import annotations

import math

import torch
from torch import nn, Tensor

from . import ops


class _GroupedLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2: Tensor, weight: Tensor, bias, groups: int):
        w2 = weight.reshape(weight.shape[0], -1).contiguous()
        ctx.save_for_backward(x2, w2)
        ctx.groups, ctx.has_bias, ctx.wshape = groups, bias is not None, weight.shape
        return ops.grouped_linear(x2, w2, bias, groups)

    @staticmethod
    def backward(ctx, gy: Tensor):
        x2, w2 = ctx.saved_tensors
        G = ctx.groups
        rows = x2.shape[0]
        Cin, Cout = w2.shape[1], w2.shape[0] // G
        gyg = gy.reshape(rows, G, Cout).permute(1, 0, 2)                  # [G, rows, Cout]
        xg = x2.reshape(rows, G, Cin).permute(1, 0, 2)                    # [G, rows, Cin]
        gw = torch.bmm(gyg.transpose(1, 2), xg).reshape(ctx.wshape)       # [G*Cout, Cin, (1)]
        gb = gy.sum(0) if ctx.has_bias else None
        gx = None
        if ctx.needs_input_grad[0]:
            gx = torch.bmm(gyg, w2.reshape(G, Cout, Cin)).permute(1, 0, 2).reshape(rows, G * Cin)
        return gx, gw, gb, None


class GroupedPointwiseConv(nn.Module):
    """``nn.Conv1d(in_channels, out_channels, kernel_size=1, groups=groups)`` over the LAST axis of
    ``[..., f]`` inputs (the reference wraps its Conv1d in two Rearranges to the same effect)."""

    def __init__(self, in_channels: int, out_channels: int, groups: int, bias: bool = True):
        super().__init__()
        if in_channels % groups or out_channels % groups:
            raise ValueError("in_channels and out_channels must be divisible by groups")
        self.in_channels, self.out_channels, self.groups = in_channels, out_channels, groups
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, 1))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):                       # nn.Conv1d's own initialisation (same RNG calls)
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in = self.weight.shape[1]
            bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x: Tensor) -> Tensor:
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        if x2.stride(-1) != 1:
            x2 = x2.contiguous()
        y = _GroupedLinearFn.apply(x2.float(), self.weight, self.bias, self.groups)
        return y.reshape(*lead, self.out_channels)
