"""SBM (PATTERN / CLUSTER) node-classification network on the fused DGN layers.

Mirrors realworld_benchmark/nets/SBMs_node_classification/dgn_net.py:8-81.
"""
import torch
import torch.nn as nn

from dgn_b200.ops import embedding
from dgn_b200.task_nets._common import build_layers
from dgn_b200.nets.mlp_readout_layer import MLPReadout


class DGNNet(nn.Module):
    def __init__(self, net_params):
        super().__init__()
        p = net_params
        self.n_classes, self.pos_enc_dim, self.device = p["n_classes"], p["pos_enc_dim"], p["device"]
        if self.pos_enc_dim > 0:
            self.embedding_pos_enc = nn.Linear(self.pos_enc_dim, p["hidden_dim"])
        self.embedding_h = nn.Embedding(p["in_dim"], p["hidden_dim"])
        self.in_feat_dropout = nn.Dropout(p["in_feat_dropout"])
        self.layers = build_layers(p)
        self.MLP_layer = MLPReadout(p["out_dim"], p["n_classes"])

    def forward(self, g, h, e, snorm_n, snorm_e):
        h = self.in_feat_dropout(embedding(self.embedding_h.weight, h, getattr(g, "n_rows_dev", None)))
        if self.pos_enc_dim > 0:
            h = h + self.embedding_pos_enc(g.ndata["pos_enc"].to(h.device))
        for conv in self.layers:
            h = conv(g, h, e, snorm_n)
        return self.MLP_layer(h)

    def loss(self, pred, label):
        # class-balanced cross-entropy: weight_c = (V - |c|) / V for classes present in the batch (:66-81)
        V = label.size(0)
        sizes = torch.bincount(label, minlength=self.n_classes)[: self.n_classes]
        weight = (V - sizes).float() / V * (sizes > 0).float()
        return nn.CrossEntropyLoss(weight=weight)(pred, label)
