"""ogbg-molpcba multi-task network on the fused DGN layers, with the optional virtual node.

Mirrors realworld_benchmark/nets/PCBA_graph_classification/dgn_net.py:9-102 (``towers`` forwarded to the layers,
``MLPReadout(out_dim, 128, decreasing_dim=...)``, ``VirtualNode`` between layers).
"""
import torch.nn as nn

from dgn_b200.compat.ogb.graphproppred.mol_encoder import AtomEncoder, BondEncoder
from dgn_b200.nets.dgn_layer import VirtualNode
from dgn_b200.nets.mlp_readout_layer import MLPReadout
from dgn_b200.task_nets._common import build_layers, graph_readout


class DGNNet(nn.Module):
    def __init__(self, net_params):
        super().__init__()
        p = net_params
        self.type_net, self.readout, self.edge_feat = p["type_net"], p["readout"], p["edge_feat"]
        self.virtual_node = p["virtual_node"]
        self.in_feat_dropout = nn.Dropout(p["in_feat_dropout"])
        self.embedding_h = AtomEncoder(emb_dim=p["hidden_dim"])
        if self.edge_feat:
            self.embedding_e = BondEncoder(emb_dim=p["edge_dim"])
        self.layers = build_layers(p, towers=p["towers"])
        self.MLP_layer = MLPReadout(p["out_dim"], 128, decreasing_dim=p["decreasing_dim"])
        self.virtual_node_layers = None
        if self.virtual_node is not None and self.virtual_node.lower() != "none":
            self.virtual_node_layers = nn.ModuleList(
                VirtualNode(dim=p["hidden_dim"], dropout=p["dropout"], batch_norm=p["batch_norm"], bias=True,
                            vn_type=self.virtual_node, residual=p["residual"]) for _ in range(p["L"] - 1))

    def forward(self, g, h, e, snorm_n, snorm_e):
        h = self.in_feat_dropout(self.embedding_h(h))
        if self.edge_feat:
            e = self.embedding_e(e)
        vn_h = 0
        for i, conv in enumerate(self.layers):
            h = conv(g, h, e, snorm_n)
            if self.virtual_node_layers is not None and i < len(self.virtual_node_layers):
                vn_h, h = self.virtual_node_layers[i](g, h, vn_h)
        return self.MLP_layer(graph_readout(g, h, self.readout if self.readout in ("sum", "max") else "mean"))

    def loss(self, scores, labels):
        return nn.BCEWithLogitsLoss()(scores, labels)
