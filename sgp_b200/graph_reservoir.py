"""DynGESN — the graph echo-state encoder, the paper's closed-form baseline.

Mirrors ``GESNLayer`` / ``GraphESN`` (lib/nn/reservoir/graph_reservoir.py:18-144) and
``GESNEncoder`` (lib/nn/encoders/dyn_gesn_encoder.py:10-44): same constructor arguments,
parameter names and RNG consumption, but the modules only hold the frozen weights; a time step

    h' = (1 - alpha) h + alpha * act( x W_ih^T + b + S (h W_hh^T) )

runs as three kernel launches per layer: ``h W_hh^T`` in the K1 scan kernel (one step, identity
activation), ``S ·`` in the K2 SpMM kernels, and ``sgp_gesn_update`` for everything else.  Unlike
the SGP encoder the propagation sits INSIDE the recurrence, so time steps cannot be batched.

Reference behaviour kept on purpose:
* ``GraphESN.__init__`` draws every layer's weights twice (each ``GESNLayer`` draws in its own
  constructor, then ``self.reset_parameters()`` redraws all, graph_reservoir.py:54,140), so the
  second pass of the RNG is what the model uses;
* ``GESNEncoder.forward`` ADDS a unit self-loop to every node (``add_self_loops`` with the node
  count inferred from ``edge_index.max() + 1``, on top of any stored diagonal) and row-normalises
  by the weighted in-degree (``normalize(dim=1)``), dyn_gesn_encoder.py:37-39; ``edge_weight=None``
  fails there with a TypeError (``None / Tensor``) and does here too;
* every layer's state is part of the output: ``[T, N, L*H]`` (``_cat_states_layers = True``).
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from . import ops
from .encoders import _opt_list, _str_to_bool
from .preprocessing import ShiftOperator, build_operator
from .reservoir import _ALLOWED, _cuda_device_for


class GESNLayer(nn.Module):
    """Frozen weights of one graph-ESN layer (graph_reservoir.py:18-93)."""

    def __init__(self, input_size, hidden_size, spectral_radius=0.9, leaking_rate=0.9, bias=True,
                 density=0.9, in_scaling=1., bias_scale=1., activation='tanh', aggr='add'):
        super().__init__()
        assert activation in _ALLOWED
        if activation == 'identity':              # tsl's get_functional_activation has no 'identity'
            raise ValueError("Activation 'identity' not valid.")
        if aggr != 'add':
            raise NotImplementedError("the reference only ever uses aggr='add' (matmul reduce)")
        self.activation_name = activation
        self.w_ih_scale, self.b_scale, self.density = in_scaling, bias_scale, density
        self.hidden_size, self.alpha, self.spectral_radius = hidden_size, leaking_rate, spectral_radius
        self.w_ih = nn.Parameter(torch.empty(hidden_size, input_size), requires_grad=False)
        self.w_hh = nn.Parameter(torch.empty(hidden_size, hidden_size), requires_grad=False)
        if bias is not None:
            self.b_ih = nn.Parameter(torch.empty(hidden_size), requires_grad=False)
        else:
            self.register_parameter('b_ih', None)
        self.reset_parameters()

    def reset_parameters(self):
        H = self.hidden_size
        self.w_ih.data.uniform_(-1, 1).mul_(self.w_ih_scale)
        if self.b_ih is not None:
            self.b_ih.data.uniform_(-1, 1).mul_(self.b_scale)
        self.w_hh.data.uniform_(-1, 1)
        if self.density < 1:
            cells = H * H
            gate = self.w_hh.data.new_ones(cells)
            gate[torch.randperm(cells)[:int(cells * (1 - self.density))]] = 0.
            self.w_hh.data.mul_(gate.view(H, H))
        rho = torch.linalg.eigvals(self.w_hh.data).abs()
        self.w_hh.data.mul_(self.spectral_radius / torch.max(rho))

    def device_weights(self, device):
        """(recurrent pack for the K1 kernel with a zero input row, zero bias, w_ih, b_ih)."""
        H = self.hidden_size
        w_hh = self.w_hh.detach().to(device=device, dtype=torch.float32)
        pack = ops.reservoir_pack(torch.zeros(H, 1, device=device), w_hh)
        w_ih = self.w_ih.detach().to(device=device, dtype=torch.float32).contiguous()
        b = None if self.b_ih is None else self.b_ih.detach().to(device=device, dtype=torch.float32).contiguous()
        return pack, torch.zeros(H, device=device), w_ih, b

    def step(self, weights, x, h_state, op: ShiftOperator, out, scratch):
        """x [N, Fin], h_state [N, H] in/out, out [N, H] view; scratch = (g_state, g, prop, zero_x)."""
        pack, zero_b, w_ih, b = weights
        g_state, g, prop, zero_x = scratch
        g_state.copy_(h_state)                    # the scan kernel overwrites its state with the result
        ops.reservoir_scan(zero_x, pack, zero_b, 1.0, "identity", g_state, g)        # g = h W_hh^T
        op.apply(g, prop)                                                             # prop = S g
        ops.gesn_update(x, w_ih, b, prop[0], self.alpha, self.activation_name, h_state, out)

    def forward(self, x, h, edge_index, edge_weight=None):
        """One step on [N, Fin] / [N, H] with an already NORMALISED adjacency (a ShiftOperator, or an
        edge list whose weights are used as they are — the reference's contract, :83-84)."""
        dev = _cuda_device_for(x)
        op = edge_index if isinstance(edge_index, ShiftOperator) else \
            _raw_operator(edge_index, edge_weight, x.size(-2), dev)
        N, H = x.size(-2), self.hidden_size
        state = h.detach().to(device=dev, dtype=torch.float32).clone().contiguous()
        xin = x.detach().to(device=dev, dtype=torch.float32).contiguous()
        out = torch.empty(N, H, device=dev)
        self.step(self.device_weights(dev), xin, state, op, out, _scratch(N, H, dev))
        return out.to(x.device)


def _scratch(N, H, dev):
    return (torch.empty(N, H, device=dev), torch.empty(1, N, H, device=dev), torch.empty(1, N, H, device=dev),
            torch.zeros(1, N, 1, device=dev))


def _raw_operator(edge_index, edge_weight, num_nodes, dev) -> ShiftOperator:
    """CSR of the given weights WITHOUT normalisation (rows = edge_index[1], duplicates kept)."""
    return build_operator(edge_index, edge_weight, num_nodes, normalize=False, device=dev)


class GraphESN(nn.Module):
    """Stack of graph-ESN layers over time (graph_reservoir.py:96-144 on tsl's _GraphRNN loop)."""

    def __init__(self, input_size, hidden_size, input_scaling=1., num_layers=1, leaking_rate=0.9,
                 spectral_radius=0.9, density=0.9, activation='tanh', bias=True, alpha_decay=False):
        super().__init__()
        self.mode, self.input_size, self.input_scaling = activation, input_size, input_scaling
        self.hidden_size, self.n_layers = hidden_size, num_layers
        self.leaking_rate, self.spectral_radius, self.density = leaking_rate, spectral_radius, density
        self.bias, self.alpha_decay = bias, alpha_decay
        cells, alpha = [], leaking_rate
        for i in range(num_layers):
            cells.append(GESNLayer(input_size=input_size if i == 0 else hidden_size, hidden_size=hidden_size,
                                   in_scaling=input_scaling, density=density, activation=activation,
                                   spectral_radius=spectral_radius, leaking_rate=alpha))
            if self.alpha_decay:
                alpha = np.clip(alpha - 0.1, 0.1, 1.)
        self.rnn_cells = nn.ModuleList(cells)
        self.reset_parameters()                   # the reference draws a second time (:140)

    def reset_parameters(self):
        for layer in self.rnn_cells:
            layer.reset_parameters()

    def forward(self, x, edge_index, edge_weight=None, h=None):
        """x [b, s, n, f] -> (out [b, s, n, L*H], h [L, b, n, H]); the graph is shared by the batch."""
        B, S, N, _ = x.shape
        dev = _cuda_device_for(x)
        op = edge_index if isinstance(edge_index, ShiftOperator) else _raw_operator(edge_index, edge_weight, N, dev)
        L, H = self.n_layers, self.hidden_size
        xd = x.detach().to(device=dev, dtype=torch.float32)
        out = torch.empty(B, S, N, L * H, device=dev)
        state = torch.zeros(L, B, N, H, device=dev) if h is None else \
            torch.as_tensor(h).detach().to(device=dev, dtype=torch.float32).clone().contiguous()
        weights = [cell.device_weights(dev) for cell in self.rnn_cells]
        scratch = _scratch(N, H, dev)
        for b in range(B):
            for s in range(S):
                inp = xd[b, s]
                for l, cell in enumerate(self.rnn_cells):
                    blk = out[b, s, :, l * H:(l + 1) * H]
                    cell.step(weights[l], inp.contiguous() if l == 0 else inp, state[l, b], op, blk, scratch)
                    inp = blk
        op.check()
        return out.to(x.device), state.to(x.device)


class GESNEncoder(nn.Module):
    """Reference: lib/nn/encoders/dyn_gesn_encoder.py:10-44."""

    def __init__(self, input_size, reservoir_size, reservoir_layers, leaking_rate, spectral_radius, density,
                 input_scaling, alpha_decay, reservoir_activation='tanh'):
        super().__init__()
        self.reservoir = GraphESN(input_size=input_size, hidden_size=reservoir_size, input_scaling=input_scaling,
                                  num_layers=reservoir_layers, leaking_rate=leaking_rate,
                                  spectral_radius=spectral_radius, density=density,
                                  activation=reservoir_activation, alpha_decay=alpha_decay)

    def forward(self, x, edge_index, edge_weight):
        """x [T, N, Fin] -> [T, N, L*H] on x's device."""
        N = x.size(-2)
        dev = _cuda_device_for(x)
        if isinstance(edge_index, ShiftOperator):          # the SparseTensor branch: used as given
            op = edge_index
        else:
            ei = torch.as_tensor(edge_index).to(device=dev, dtype=torch.int64)
            if edge_weight is None:                        # normalize(): `None / degree[index]`
                raise TypeError("unsupported operand type(s) for /: 'NoneType' and 'Tensor'")
            ew = torch.as_tensor(edge_weight).to(device=dev, dtype=torch.float32)
            n_loops = int(ei.max()) + 1 if ei.numel() else 0        # add_self_loops infers N from the edges
            loops = torch.arange(n_loops, device=dev, dtype=torch.int64)
            ei = torch.cat([ei, torch.stack([loops, loops])], 1)
            ew = torch.cat([ew, torch.ones(n_loops, device=dev)])
            op = build_operator(ei, ew, N, device=dev)     # e_ij / weighted in-degree of the target row
        out, _ = self.reservoir(x[None], op)
        return out[0]

    @staticmethod
    def add_model_specific_args(parser):
        _opt_list(parser, '--reservoir-size', type=int, default=32, tunable=True, options=[16, 32, 64, 128, 256])
        _opt_list(parser, '--reservoir-layers', type=int, default=1, tunable=True, options=[1, 2, 3])
        _opt_list(parser, '--spectral-radius', type=float, default=0.9, tunable=True, options=[0.7, 0.8, 0.9])
        _opt_list(parser, '--leaking-rate', type=float, default=0.9, tunable=True, options=[0.7, 0.8, 0.9])
        _opt_list(parser, '--density', type=float, default=0.7, tunable=True, options=[0.7, 0.8, 0.9])
        _opt_list(parser, '--input-scaling', type=float, default=1., tunable=True, options=[1., 1.5, 2.])
        parser.add_argument('--reservoir-activation', type=str, default='tanh')
        _opt_list(parser, '--alpha-decay', type=_str_to_bool, nargs='?', const=True, default=False)
        return parser
