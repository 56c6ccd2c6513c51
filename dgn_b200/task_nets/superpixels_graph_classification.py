"""Superpixel (CIFAR10 / MNIST) graph-classification network on the fused DGN layers.

Mirrors realworld_benchmark/nets/superpixels_graph_classification/dgn_net.py:7-78.
"""
import torch.nn as nn

from dgn_b200.task_nets._common import build_layers, graph_readout
from dgn_b200.nets.mlp_readout_layer import MLPReadout


class DGNNet(nn.Module):
    def __init__(self, net_params):
        super().__init__()
        p = net_params
        self.readout, self.edge_feat = p["readout"], p["edge_feat"]
        self.embedding_h = nn.Linear(p["in_dim"], p["hidden_dim"])
        self.in_feat_dropout = nn.Dropout(p["in_feat_dropout"])
        if self.edge_feat:
            self.embedding_e = nn.Linear(p["in_dim_edge"], p["edge_dim"])
        self.layers = build_layers(p)
        self.MLP_layer = MLPReadout(p["out_dim"], p["n_classes"])

    def forward(self, g, h, e, snorm_n, snorm_e):
        h = self.in_feat_dropout(self.embedding_h(h))
        if self.edge_feat:
            e = self.embedding_e(e)
        for conv in self.layers:
            h = conv(g, h, e, snorm_n)
        return self.MLP_layer(graph_readout(g, h, self.readout if self.readout in ("sum", "max") else "mean"))

    def loss(self, pred, label):
        return nn.CrossEntropyLoss()(pred, label)
