"""Graph-level prediction head: ``FC_layers.{l}`` as in realworld_benchmark/nets/mlp_readout_layer.py:11-30."""
import torch
import torch.nn as nn


class MLPReadout(nn.Module):
    def __init__(self, input_dim, output_dim, L=2, decreasing_dim=True):
        super().__init__()
        widths = [input_dim // 2 ** l if decreasing_dim else input_dim for l in range(L + 1)] + [output_dim]
        self.FC_layers = nn.ModuleList(nn.Linear(widths[l], widths[l + 1], bias=True) for l in range(L + 1))
        self.L = L

    def forward(self, x):
        if x.is_cuda:
            from dgn_b200 import ops
            if ops.head_supported(x, self.FC_layers):          # L = 2 on a batch of graph vectors: one launch
                return ops.mlp_head(x, self.FC_layers)
        for fc in self.FC_layers[:-1]:
            x = torch.relu(fc(x))
        return self.FC_layers[-1](x)
