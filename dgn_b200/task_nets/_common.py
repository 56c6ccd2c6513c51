"""Pieces shared by the task networks: the DGN layer stack and the graph readouts."""
from __future__ import annotations

import torch
import torch.nn as nn

from dgn_b200.ops import readout

from dgn_b200.nets.dgn_layer import DGNLayer


def build_layers(net_params, **extra):
    p = net_params
    dims = [p["hidden_dim"]] * p["L"] + [p["out_dim"]]
    return nn.ModuleList(
        DGNLayer(in_dim=dims[i], out_dim=dims[i + 1], dropout=p["dropout"], graph_norm=p["graph_norm"],
                 batch_norm=p["batch_norm"], residual=p["residual"], aggregators=p["aggregators"],
                 scalers=p["scalers"], avg_d=p["avg_d"], type_net=p["type_net"], edge_features=p["edge_feat"],
                 edge_dim=p["edge_dim"], pretrans_layers=p["pretrans_layers"],
                 posttrans_layers=p["posttrans_layers"], **extra).model
        for i in range(p["L"]))


def graph_readout(g, h, mode):
    """dgl.{sum,max,mean}_nodes and the two directional readouts of the ZINC net (dgn_net.py:71-86)."""
    if mode in ("sum", "max"):
        return readout(g, h, mode)
    if mode in ("directional", "directional_abs"):
        e1 = g.ndata["eig"][:, 1:2].to(h.device)
        # the reference divides by sum(|eig_1|, dim=1) of a single column, i.e. by |eig_1| itself
        denom = torch.sum(torch.abs(e1), dim=1, keepdim=True)
        if mode == "directional_abs":
            return torch.cat([readout(g, h * torch.abs(e1) / denom, "mean"), readout(g, h, "mean")], dim=1)
        return torch.cat([torch.abs(readout(g, h * e1 / denom, "mean")), readout(g, h, "mean")], dim=1)
    return readout(g, h, "mean")
