"""FCLayer / MLP with the reference's constructor, parameter names and init.

Mirrors realworld_benchmark/nets/layers.py:21-154 (``fully_connected.{i}.linear.{weight,bias}``;
``xavier_uniform_(weight, gain=1/in_size)``, zero bias; torch RNG consumed in the same order).
These are the "small dense GEMMs" of the north star: plain library GEMMs in fp32.
"""
from __future__ import annotations

import torch
import torch.nn as nn

SUPPORTED_ACTIVATION_MAP = {"ReLU", "Sigmoid", "Tanh", "ELU", "SELU", "GLU", "LeakyReLU", "Softplus", "None"}


def get_activation(activation):
    if activation and callable(activation):
        return activation
    match = [x for x in SUPPORTED_ACTIVATION_MAP if str(activation).lower() == x.lower()]
    assert len(match) == 1, "Unhandled activation function"
    return None if match[0] == "None" else getattr(torch.nn.modules.activation, match[0])()


class FCLayer(nn.Module):
    def __init__(self, in_size, out_size, activation="relu", dropout=0., b_norm=False, bias=True, init_fn=None,
                 device="cpu"):
        super().__init__()
        self.in_size, self.out_size, self.bias = in_size, out_size, bias
        self.linear = nn.Linear(in_size, out_size, bias=bias).to(device)
        # the reference builds nn.Dropout(p, device=...) which raises for dropout > 0 (layers.py:86);
        # a working Dropout is the intended behaviour
        self.dropout = nn.Dropout(p=dropout) if dropout else None
        self.b_norm = nn.BatchNorm1d(out_size).to(device) if b_norm else None
        self.activation = get_activation(activation)
        self.init_fn = nn.init.xavier_uniform_
        self.reset_parameters()

    def reset_parameters(self, init_fn=None):
        init_fn = init_fn or self.init_fn
        if init_fn is not None:
            init_fn(self.linear.weight, 1 / self.in_size)       # second positional arg is the GAIN
        if self.bias:
            self.linear.bias.data.zero_()

    def forward(self, x):
        h = self.linear(x)
        if self.activation is not None:
            h = self.activation(h)
        if self.dropout is not None:
            h = self.dropout(h)
        if self.b_norm is not None:
            h = self.b_norm(h.transpose(1, 2)).transpose(1, 2) if h.shape[1] != self.out_size else self.b_norm(h)
        return h

    def __repr__(self):
        return "%s (%d -> %d)" % (self.__class__.__name__, self.in_size, self.out_size)


class MLP(nn.Module):
    """``layers`` FCLayers: in -> hidden -> ... -> out; a single layer maps in -> out directly."""

    def __init__(self, in_size, hidden_size, out_size, layers, mid_activation="relu", last_activation="none",
                 dropout=0., mid_b_norm=False, last_b_norm=False, device="cpu"):
        super().__init__()
        self.in_size, self.hidden_size, self.out_size = in_size, hidden_size, out_size
        depth = max(int(layers), 1)
        self.fully_connected = nn.ModuleList()
        for i in range(depth):
            last = i == depth - 1
            self.fully_connected.append(
                FCLayer(in_size if i == 0 else hidden_size, out_size if last else hidden_size,
                        activation=last_activation if last else mid_activation,
                        b_norm=last_b_norm if last else mid_b_norm, device=device, dropout=dropout))

    def forward(self, x):
        for fc in self.fully_connected:
            x = fc(x)
        return x

    def __repr__(self):
        return "%s (%d -> %d)" % (self.__class__.__name__, self.in_size, self.out_size)
