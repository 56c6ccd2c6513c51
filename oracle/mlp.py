"""FCLayer / MLP of the reference restated (realworld_benchmark/nets/layers.py:21-154).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Sub-module names (``linear``, ``fully_connected``)
and the order in which the torch RNG is consumed are those of the reference, so a
``state_dict`` - or just ``torch.manual_seed`` - interchanges between the reference, this
oracle and ``dgn_b200.nets``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

_ACTIVATIONS = ("ReLU", "Sigmoid", "Tanh", "ELU", "SELU", "GLU", "LeakyReLU", "Softplus", "None")


def resolve_activation(spec):
    """layers.py:7-18: a callable passes through, a (case-insensitive) name maps to ``torch.nn``."""
    if spec and callable(spec):
        return spec
    hits = [a for a in _ACTIVATIONS if a.lower() == str(spec).lower()]
    assert len(hits) == 1, "Unhandled activation function"
    return None if hits[0] == "None" else getattr(nn, hits[0])()


class FCLayer(nn.Module):
    """Linear -> activation -> dropout -> batch-norm (layers.py:76-111).

    Init (layers.py:94-99): ``xavier_uniform_(weight, gain=1/in_size)`` - the reference passes
    ``1/in_size`` positionally, which torch reads as the *gain* - and a zero bias.
    """

    def __init__(self, in_size, out_size, activation="relu", dropout=0.0, b_norm=False, bias=True):
        super().__init__()
        self.in_size, self.out_size, self.bias = in_size, out_size, bias
        self.linear = nn.Linear(in_size, out_size, bias=bias)
        self.dropout = nn.Dropout(p=dropout) if dropout else None
        self.b_norm = nn.BatchNorm1d(out_size) if b_norm else None
        self.activation = resolve_activation(activation)
        nn.init.xavier_uniform_(self.linear.weight, 1 / in_size)
        if bias:
            self.linear.bias.data.zero_()

    def forward(self, x):
        y = self.linear(x)
        if self.activation is not None:
            y = self.activation(y)
        if self.dropout is not None:
            y = self.dropout(y)
        if self.b_norm is not None:
            y = self.b_norm(y)
        return y


class MLP(nn.Module):
    """Stack of FCLayers (layers.py:125-149): ``layers<=1`` is a single Linear with ``last_activation``."""

    def __init__(self, in_size, hidden_size, out_size, layers, mid_activation="relu", last_activation="none"):
        super().__init__()
        widths = [in_size] + [hidden_size] * (max(layers, 1) - 1) + [out_size]
        acts = [mid_activation] * (len(widths) - 2) + [last_activation]
        self.fully_connected = nn.ModuleList(
            FCLayer(widths[i], widths[i + 1], activation=acts[i]) for i in range(len(acts)))

    def forward(self, x):
        for fc in self.fully_connected:
            x = fc(x)
        return x


class MLPReadout(nn.Module):
    """realworld_benchmark/nets/mlp_readout_layer.py:11-30: L halving Linear+ReLU, then Linear."""

    def __init__(self, input_dim, output_dim, L=2, decreasing_dim=True):
        super().__init__()
        dims = [input_dim // 2 ** l if decreasing_dim else input_dim for l in range(L + 1)]     # :13-18
        self.FC_layers = nn.ModuleList(
            [nn.Linear(dims[l], dims[l + 1], bias=True) for l in range(L)] + [nn.Linear(dims[L], output_dim, bias=True)])
        self.L = L

    def forward(self, x):
        for l in range(self.L):
            x = torch.relu(self.FC_layers[l](x))
        return self.FC_layers[self.L](x)
