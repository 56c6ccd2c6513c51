"""ZINC graph-regression network on the fused DGN layers.

Same ``net_params`` keys, parameter names and ``forward(g, h, e, snorm_n, snorm_e)`` / ``loss`` as
realworld_benchmark/nets/molecules_graph_regression/dgn_net.py:8-92 (the reference file itself also
runs unmodified on top of ``nets.dgn_layer`` from this package).
"""
import torch.nn as nn

from dgn_b200.ops import embedding, l1_loss
from dgn_b200.task_nets._common import build_layers, graph_readout
from dgn_b200.nets.mlp_readout_layer import MLPReadout


class DGNNet(nn.Module):
    def __init__(self, net_params):
        super().__init__()
        p = net_params
        self.type_net, self.pos_enc_dim, self.readout = p["type_net"], p["pos_enc_dim"], p["readout"]
        self.edge_feat, self.device = p["edge_feat"], p["device"]
        if self.pos_enc_dim > 0:
            self.embedding_pos_enc = nn.Linear(self.pos_enc_dim, p["hidden_dim"])
        self.in_feat_dropout = nn.Dropout(p["in_feat_dropout"])
        self.embedding_h = nn.Embedding(p["num_atom_type"], p["hidden_dim"])
        if self.edge_feat:
            self.embedding_e = nn.Embedding(p["num_bond_type"], p["edge_dim"])
        self.layers = build_layers(p)
        wide = self.readout in ("directional", "directional_abs")
        self.MLP_layer = MLPReadout((2 if wide else 1) * p["out_dim"], 1)

    def forward(self, g, h, e, snorm_n, snorm_e):
        h = self.in_feat_dropout(embedding(self.embedding_h.weight, h, getattr(g, "n_rows_dev", None)))
        if self.pos_enc_dim > 0:
            h = h + self.embedding_pos_enc(g.ndata["pos_enc"].to(h.device))
        if self.edge_feat:
            e = self.embedding_e(e)
        for conv in self.layers:
            h = conv(g, h, e, snorm_n)
        return self.MLP_layer(graph_readout(g, h, self.readout))

    def loss(self, scores, targets):
        return l1_loss(scores, targets)          # nn.L1Loss()(scores, targets), one launch per direction on CUDA
