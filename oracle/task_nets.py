"""Task networks that call the DGN layers, restated (TEST INFRASTRUCTURE, see oracle/__init__.py).

* ``ZincNet``        realworld_benchmark/nets/molecules_graph_regression/dgn_net.py:8-92
* ``PatternNet``     realworld_benchmark/nets/SBMs_node_classification/dgn_net.py:8-81
* ``SuperpixelNet``  realworld_benchmark/nets/superpixels_graph_classification/dgn_net.py:7-78
* ``HivNet``         realworld_benchmark/nets/HIV_graph_classification/dgn_net.py:13-88
* ``PcbaNet``        realworld_benchmark/nets/PCBA_graph_classification/dgn_net.py:9-102

Parameter names (``embedding_h``, ``layers.{i}.*``, ``MLP_layer.FC_layers.{l}``) and module
construction order follow the reference so seeds and state_dicts interchange.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .directional_layers import DGNLayer, VirtualNode
from .mlp import MLPReadout
from . import use_standin_dgl


def _conv_stack(p):
    dims = [p["hidden_dim"]] * p["L"] + [p["out_dim"]]
    return nn.ModuleList(
        DGNLayer(in_dim=dims[i], out_dim=dims[i + 1], dropout=p["dropout"], graph_norm=p["graph_norm"],
                 batch_norm=p["batch_norm"], residual=p["residual"], aggregators=p["aggregators"],
                 scalers=p["scalers"], avg_d=p["avg_d"], type_net=p["type_net"], edge_features=p["edge_feat"],
                 edge_dim=p["edge_dim"], pretrans_layers=p["pretrans_layers"],
                 posttrans_layers=p["posttrans_layers"], **p.get("layer_kwargs", {})).model
        for i in range(p["L"]))


def _graph_readout(g, h, mode):
    dgl = use_standin_dgl()
    g.ndata["h"] = h
    if mode == "sum":
        return dgl.sum_nodes(g, "h")
    if mode == "max":
        return dgl.max_nodes(g, "h")
    if mode in ("directional", "directional_abs"):        # molecules dgn_net.py:77-84
        e1 = g.ndata["eig"][:, 1:2]
        if mode == "directional_abs":
            g.ndata["dir"] = h * torch.abs(e1) / torch.sum(torch.abs(e1), dim=1, keepdim=True)
            return torch.cat([dgl.mean_nodes(g, "dir"), dgl.mean_nodes(g, "h")], dim=1)
        g.ndata["dir"] = h * e1 / torch.sum(torch.abs(e1), dim=1, keepdim=True)
        return torch.cat([torch.abs(dgl.mean_nodes(g, "dir")), dgl.mean_nodes(g, "h")], dim=1)
    return dgl.mean_nodes(g, "h")


class ZincNet(nn.Module):
    def __init__(self, net_params):
        super().__init__()
        p = net_params
        self.pos_enc_dim, self.readout, self.edge_feat = p["pos_enc_dim"], p["readout"], p["edge_feat"]
        if self.pos_enc_dim > 0:
            self.embedding_pos_enc = nn.Linear(self.pos_enc_dim, p["hidden_dim"])
        self.in_feat_dropout = nn.Dropout(p["in_feat_dropout"])
        self.embedding_h = nn.Embedding(p["num_atom_type"], p["hidden_dim"])
        if self.edge_feat:
            self.embedding_e = nn.Embedding(p["num_bond_type"], p["edge_dim"])
        self.layers = _conv_stack(p)
        wide = self.readout in ("directional", "directional_abs")
        self.MLP_layer = MLPReadout((2 if wide else 1) * p["out_dim"], 1)

    def forward(self, g, h, e, snorm_n, snorm_e):           # dgn_net.py:57-88
        h = self.in_feat_dropout(self.embedding_h(h))
        if self.pos_enc_dim > 0:
            h = h + self.embedding_pos_enc(g.ndata["pos_enc"])
        if self.edge_feat:
            e = self.embedding_e(e)
        for conv in self.layers:
            h = conv(g, h, e, snorm_n)
        return self.MLP_layer(_graph_readout(g, h, self.readout))

    def loss(self, scores, targets):                        # :90-92
        return nn.L1Loss()(scores, targets)


class PatternNet(nn.Module):
    def __init__(self, net_params):
        super().__init__()
        p = net_params
        self.n_classes, self.pos_enc_dim = p["n_classes"], p["pos_enc_dim"]
        if self.pos_enc_dim > 0:
            self.embedding_pos_enc = nn.Linear(self.pos_enc_dim, p["hidden_dim"])
        self.embedding_h = nn.Embedding(p["in_dim"], p["hidden_dim"])
        self.in_feat_dropout = nn.Dropout(p["in_feat_dropout"])
        self.layers = _conv_stack(p)
        self.MLP_layer = MLPReadout(p["out_dim"], p["n_classes"])

    def forward(self, g, h, e, snorm_n, snorm_e):           # SBMs dgn_net.py:53-64
        h = self.in_feat_dropout(self.embedding_h(h))
        if self.pos_enc_dim > 0:
            h = h + self.embedding_pos_enc(g.ndata["pos_enc"])
        for conv in self.layers:
            h = conv(g, h, e, snorm_n)
        return self.MLP_layer(h)

    def loss(self, pred, label):                            # :66-81  class-balanced cross-entropy
        V = label.size(0)
        sizes = torch.bincount(label, minlength=self.n_classes)[: self.n_classes]
        weight = (V - sizes).float() / V
        weight = weight * (sizes > 0).float()
        return nn.CrossEntropyLoss(weight=weight)(pred, label)


class SuperpixelNet(nn.Module):
    def __init__(self, net_params):
        super().__init__()
        p = net_params
        self.readout, self.edge_feat = p["readout"], p["edge_feat"]
        self.embedding_h = nn.Linear(p["in_dim"], p["hidden_dim"])
        self.in_feat_dropout = nn.Dropout(p["in_feat_dropout"])
        if self.edge_feat:
            self.embedding_e = nn.Linear(p["in_dim_edge"], p["edge_dim"])
        self.layers = _conv_stack(p)
        self.MLP_layer = MLPReadout(p["out_dim"], p["n_classes"])

    def forward(self, g, h, e, snorm_n, snorm_e):           # superpixels dgn_net.py:52-73
        h = self.in_feat_dropout(self.embedding_h(h))
        if self.edge_feat:
            e = self.embedding_e(e)
        for conv in self.layers:
            h = conv(g, h, e, snorm_n)
        return self.MLP_layer(_graph_readout(g, h, self.readout))

    def loss(self, pred, label):                            # :75-78
        return nn.CrossEntropyLoss()(pred, label)


def _mol_encoders():
    use_standin_dgl()                       # puts oracle/standin (dgl AND the ogb stand-in) on sys.path
    from ogb.graphproppred.mol_encoder import AtomEncoder, BondEncoder
    return AtomEncoder, BondEncoder


class HivNet(nn.Module):
    def __init__(self, net_params):
        super().__init__()
        p = net_params
        AtomEncoder, BondEncoder = _mol_encoders()
        self.pos_enc_dim, self.readout, self.edge_feat = p["pos_enc_dim"], p["readout"], p["edge_feat"]
        if self.pos_enc_dim > 0:
            self.embedding_pos_enc = nn.Linear(self.pos_enc_dim, p["hidden_dim"])
        self.in_feat_dropout = nn.Dropout(p["in_feat_dropout"])
        self.embedding_h = AtomEncoder(emb_dim=p["hidden_dim"])
        if self.edge_feat:
            self.embedding_e = BondEncoder(emb_dim=p["edge_dim"])
        self.layers = _conv_stack(p)         # the reference does not forward `towers` here (layer default: 5)
        self.MLP_layer = MLPReadout(p["out_dim"], 1)

    def forward(self, g, h, e, snorm_n, snorm_e):           # HIV dgn_net.py:62-84
        h = self.in_feat_dropout(self.embedding_h(h))
        if self.pos_enc_dim > 0:
            h = h + self.embedding_pos_enc(g.ndata["pos_enc"])
        if self.edge_feat:
            e = self.embedding_e(e)
        for conv in self.layers:
            h = conv(g, h, e, snorm_n)
        return self.MLP_layer(_graph_readout(g, h, self.readout if self.readout in ("sum", "max") else "mean"))

    def loss(self, scores, labels):                         # :86-88 (minus the hard-coded .to('cuda'))
        return nn.BCEWithLogitsLoss()(scores, labels.float().unsqueeze(-1))


class PcbaNet(nn.Module):
    def __init__(self, net_params):
        super().__init__()
        p = net_params
        AtomEncoder, BondEncoder = _mol_encoders()
        self.readout, self.edge_feat, self.virtual_node = p["readout"], p["edge_feat"], p["virtual_node"]
        self.in_feat_dropout = nn.Dropout(p["in_feat_dropout"])
        self.embedding_h = AtomEncoder(emb_dim=p["hidden_dim"])
        if self.edge_feat:
            self.embedding_e = BondEncoder(emb_dim=p["edge_dim"])
        self.layers = _conv_stack(dict(p, layer_kwargs={"towers": p["towers"]}))
        self.MLP_layer = MLPReadout(p["out_dim"], 128, decreasing_dim=p["decreasing_dim"])
        self.virtual_node_layers = None
        if self.virtual_node is not None and self.virtual_node.lower() != "none":       # :60-66
            self.virtual_node_layers = nn.ModuleList(
                VirtualNode(dim=p["hidden_dim"], dropout=p["dropout"], batch_norm=p["batch_norm"], bias=True,
                            vn_type=self.virtual_node, residual=p["residual"]) for _ in range(p["L"] - 1))

    def forward(self, g, h, e, snorm_n, snorm_e):           # PCBA dgn_net.py:68-97
        h = self.in_feat_dropout(self.embedding_h(h))
        if self.edge_feat:
            e = self.embedding_e(e)
        vn_h = 0
        for i, conv in enumerate(self.layers):
            h = conv(g, h, e, snorm_n)
            if self.virtual_node_layers is not None and i < len(self.virtual_node_layers):
                vn_h, h = self.virtual_node_layers[i](g, h, vn_h)
        return self.MLP_layer(_graph_readout(g, h, self.readout if self.readout in ("sum", "max") else "mean"))

    def loss(self, scores, labels):                         # :99-102
        return nn.BCEWithLogitsLoss()(scores, labels)
