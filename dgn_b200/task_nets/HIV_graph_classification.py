"""ogbg-molhiv graph-classification network on the fused DGN layers.

Same ``net_params`` keys, parameter names and ``forward(g, h, e, snorm_n, snorm_e)`` / ``loss`` as
realworld_benchmark/nets/HIV_graph_classification/dgn_net.py:13-88.  The reference never forwards ``towers`` to
``DGNLayer`` (its tower count is the layer default, 5); an optional ``net_params['towers']`` is honoured here the way
the PCBA net does it (rb/nets/PCBA_graph_classification/dgn_net.py:46,54) so that BASELINE configs[3] (4 towers) is
expressible.
"""
import torch
import torch.nn as nn

from dgn_b200.compat.ogb.graphproppred.mol_encoder import AtomEncoder, BondEncoder
from dgn_b200.nets.mlp_readout_layer import MLPReadout
from dgn_b200.task_nets._common import build_layers, graph_readout


class DGNNet(nn.Module):
    def __init__(self, net_params):
        super().__init__()
        p = net_params
        self.type_net, self.pos_enc_dim, self.readout = p["type_net"], p["pos_enc_dim"], p["readout"]
        self.edge_feat, self.device = p["edge_feat"], p["device"]
        if self.pos_enc_dim > 0:
            self.embedding_pos_enc = nn.Linear(self.pos_enc_dim, p["hidden_dim"])
        self.in_feat_dropout = nn.Dropout(p["in_feat_dropout"])
        self.embedding_h = AtomEncoder(emb_dim=p["hidden_dim"])
        if self.edge_feat:
            self.embedding_e = BondEncoder(emb_dim=p["edge_dim"])
        extra = {"towers": p["towers"]} if "towers" in p else {}
        self.layers = build_layers(p, **extra)
        self.MLP_layer = MLPReadout(p["out_dim"], 1)

    def forward(self, g, h, e, snorm_n, snorm_e):
        h = self.in_feat_dropout(self.embedding_h(h))
        if self.pos_enc_dim > 0:
            h = h + self.embedding_pos_enc(g.ndata["pos_enc"].to(h.device))
        if self.edge_feat:
            e = self.embedding_e(e)
        for conv in self.layers:
            h = conv(g, h, e, snorm_n)
        return self.MLP_layer(graph_readout(g, h, self.readout if self.readout in ("sum", "max") else "mean"))

    def loss(self, scores, labels):
        # the reference moves the labels to 'cuda' unconditionally (:88); here they follow the scores
        return nn.BCEWithLogitsLoss()(scores, labels.to(scores.device, torch.float32).unsqueeze(-1))
