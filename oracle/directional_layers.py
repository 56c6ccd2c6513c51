"""DGN layers of the reference restated on the degree-bucketed (DGL-style) graph API.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows realworld_benchmark/nets/dgn_layer.py:

* ``ComplexConv``  = DGNLayerComplex  :52-132
* ``SimpleConv``   = DGNLayerSimple   :135-202
* ``TowerConv``    = DGNTower         :205-276   (no ReLU, no residual)
* ``TowerStack``   = DGNLayerTower    :279-325
* ``DGNLayer``     = factory          :328-352   (callers use ``.model``)
* ``VirtualNode``  = VirtualNode      :12-49     (PCBA net only)

It runs the same execution structure as the reference - ``apply_edges`` with a pretrans UDF,
then ``update_all`` whose reduce UDF is called once per distinct in-degree - which is why it
also serves as the timed CPU baseline ("port" of the reference's python path).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .mailbox_ops import AGGREGATORS, SCALERS, reduce_bucket
from .mlp import MLP, FCLayer


class _BucketedConv(nn.Module):
    """What the three reference layer classes share: message UDFs + the bucketed reduce."""

    edge_pretrans = True        # complex / tower: message = pretrans(cat(h_u, h_v[, ef]))

    def _setup(self, in_dim, out_dim, dropout, graph_norm, batch_norm, aggregators, scalers, avg_d,
               edge_features=False, edge_dim=0, pretrans_layers=1, posttrans_layers=1, concat_input=True):
        self.dropout, self.graph_norm, self.batch_norm = dropout, graph_norm, batch_norm
        self.edge_features = bool(edge_features)
        self.aggregators, self.scalers, self.avg_d = aggregators, scalers, avg_d
        self.batchnorm_h = nn.BatchNorm1d(out_dim)
        if self.edge_pretrans:
            self.pretrans = MLP(2 * in_dim + (edge_dim if edge_features else 0), in_dim, in_dim, pretrans_layers)
        width = len(aggregators) * len(scalers) + (1 if concat_input else 0)
        self.posttrans = MLP(width * in_dim, out_dim, out_dim, posttrans_layers)

    # dgn_layer.py:75-84 / :154-159
    def _edge_udf(self, edges):
        if not self.edge_pretrans:
            msg = edges.src["h"]
        else:
            parts = [edges.src["h"], edges.dst["h"]] + ([edges.data["ef"]] if self.edge_features else [])
            msg = self.pretrans(torch.cat(parts, dim=1))
        return {"e": msg, "eig_s": edges.src["eig"], "eig_d": edges.dst["eig"]}

    @staticmethod
    def _message_udf(edges):
        return {"e": edges.data["e"], "eig_s": edges.data["eig_s"], "eig_d": edges.data["eig_d"]}

    # dgn_layer.py:86-98
    def _reduce_udf(self, nodes):
        box = nodes.mailbox
        return {"h": reduce_bucket(box["e"], box["eig_s"], box["eig_d"], nodes.data["h"],
                                   self.aggregators, self.scalers, self.avg_d)}

    def _aggregate(self, g, h, e):
        g.ndata["h"] = h
        if self.edge_pretrans and self.edge_features:
            g.edata["ef"] = e
        g.apply_edges(self._edge_udf)
        g.update_all(self._message_udf, self._reduce_udf)
        return g.ndata["h"]

    def _normalise(self, h, snorm_n):
        if self.graph_norm:
            h = h * snorm_n
        if self.batch_norm:
            h = self.batchnorm_h(h)
        return h


class ComplexConv(_BucketedConv):
    def __init__(self, in_dim, out_dim, dropout, graph_norm, batch_norm, aggregators, scalers, avg_d, residual,
                 edge_features, edge_dim, pretrans_layers=1, posttrans_layers=1):
        super().__init__()
        self._setup(in_dim, out_dim, dropout, graph_norm, batch_norm, aggregators, scalers, avg_d,
                    edge_features, edge_dim, pretrans_layers, posttrans_layers, concat_input=True)
        self.residual = residual and in_dim == out_dim          # :72-73

    def forward(self, g, h, e, snorm_n):                        # :103-132
        h_in = h
        h = self.posttrans(torch.cat([h, self._aggregate(g, h, e)], dim=1))
        h = F.relu(self._normalise(h, snorm_n))
        if self.residual:
            h = h_in + h
        return F.dropout(h, self.dropout, training=self.training)


class SimpleConv(_BucketedConv):
    edge_pretrans = False

    def __init__(self, in_dim, out_dim, dropout, graph_norm, batch_norm, aggregators, scalers, residual, avg_d,
                 posttrans_layers=1):
        super().__init__()
        self._setup(in_dim, out_dim, dropout, graph_norm, batch_norm, aggregators, scalers, avg_d,
                    posttrans_layers=posttrans_layers, concat_input=False)
        self.residual = residual and in_dim == out_dim          # :151-152

    def forward(self, g, h, e, snorm_n):                        # :178-202
        h_in = h
        h = self.posttrans(self._aggregate(g, h, e))
        h = F.relu(self._normalise(h, snorm_n))
        if self.residual:
            h = h_in + h
        return F.dropout(h, self.dropout, training=self.training)


class TowerConv(_BucketedConv):
    def __init__(self, in_dim, out_dim, dropout, graph_norm, batch_norm, aggregators, scalers, avg_d,
                 pretrans_layers, posttrans_layers, edge_features, edge_dim):
        super().__init__()
        self._setup(in_dim, out_dim, dropout, graph_norm, batch_norm, aggregators, scalers, avg_d,
                    edge_features, edge_dim, pretrans_layers, posttrans_layers, concat_input=True)

    def forward(self, g, h, e, snorm_n):                        # :254-276
        h = self.posttrans(torch.cat([h, self._aggregate(g, h, e)], dim=1))
        h = self._normalise(h, snorm_n)
        return F.dropout(h, self.dropout, training=self.training)


class TowerStack(nn.Module):
    def __init__(self, in_dim, out_dim, aggregators, scalers, avg_d, dropout, graph_norm, batch_norm, towers=5,
                 pretrans_layers=1, posttrans_layers=1, divide_input=True, residual=False, edge_features=False,
                 edge_dim=0):
        super().__init__()
        assert (not divide_input) or in_dim % towers == 0, "towers must divide in_dim when divide_input is set"
        assert out_dim % towers == 0, "towers must divide out_dim"
        assert avg_d is not None
        self.divide_input = divide_input
        self.input_tower = in_dim // towers if divide_input else in_dim
        self.output_tower = out_dim // towers
        self.residual = residual and in_dim == out_dim          # :297-298
        self.towers = nn.ModuleList(
            TowerConv(self.input_tower, self.output_tower, dropout, graph_norm, batch_norm, aggregators, scalers,
                      avg_d, pretrans_layers, posttrans_layers, edge_features, edge_dim) for _ in range(towers))
        self.mixing_network = FCLayer(out_dim, out_dim, activation="LeakyReLU")

    def forward(self, g, h, e, snorm_n):                        # :309-325
        w = self.input_tower
        outs = [tw(g, h[:, i * w:(i + 1) * w] if self.divide_input else h, e, snorm_n)
                for i, tw in enumerate(self.towers)]
        y = torch.cat(outs, dim=1)
        if len(self.towers) > 1:
            y = self.mixing_network(y)
        return h + y if self.residual else y


class DGNLayer(nn.Module):
    """Factory of dgn_layer.py:328-352: resolves the registry names and exposes ``.model``."""

    def __init__(self, in_dim, out_dim, dropout, graph_norm, batch_norm, aggregators, scalers, avg_d, type_net,
                 residual, towers=5, divide_input=True, edge_features=None, edge_dim=None, pretrans_layers=1,
                 posttrans_layers=1):
        super().__init__()
        aggs = [AGGREGATORS[a] for a in aggregators.split()]
        scs = [SCALERS[s] for s in scalers.split()]
        if type_net == "simple":
            self.model = SimpleConv(in_dim, out_dim, dropout, graph_norm, batch_norm, aggs, scs, residual, avg_d,
                                    posttrans_layers)
        elif type_net == "complex":
            self.model = ComplexConv(in_dim, out_dim, dropout, graph_norm, batch_norm, aggs, scs, avg_d, residual,
                                     edge_features, edge_dim, pretrans_layers, posttrans_layers)
        elif type_net == "towers":
            self.model = TowerStack(in_dim, out_dim, aggs, scs, avg_d, dropout, graph_norm, batch_norm, towers,
                                    pretrans_layers, posttrans_layers, divide_input, residual, edge_features,
                                    edge_dim)


class VirtualNode(nn.Module):
    """realworld_benchmark/nets/dgn_layer.py:12-49: per-graph pooled feature pushed through an FCLayer and added back
    to every node of its graph."""

    def __init__(self, dim, dropout, batch_norm=False, bias=True, residual=True, vn_type="mean"):
        super().__init__()
        self.vn_type = vn_type.lower()
        self.fc_layer = FCLayer(in_size=dim, out_size=dim, activation="relu", dropout=dropout, b_norm=batch_norm,
                                bias=bias)
        self.residual = residual

    def forward(self, g, h, vn_h):
        from . import use_standin_dgl
        dgl = use_standin_dgl()
        g.ndata["h"] = h
        if self.vn_type == "mean":                                        # :25-33
            pool = dgl.mean_nodes(g, "h")
        elif self.vn_type == "sum":
            pool = dgl.sum_nodes(g, "h")
        elif self.vn_type == "logsum":
            lognum = torch.log(torch.tensor(g.batch_num_nodes, dtype=h.dtype, device=h.device))
            pool = dgl.mean_nodes(g, "h") * lognum.unsqueeze(-1)
        else:
            raise ValueError("Undefined input %r. Accepted values are sum, mean, logsum" % self.vn_type)
        vn_new = self.fc_layer(vn_h + pool)                               # :37-41
        vn_h = vn_h + vn_new if self.residual else vn_new
        spread = torch.cat([vn_h[i:i + 1].repeat(n, 1) for i, n in enumerate(g.batch_num_nodes)], dim=0)   # :44-46
        return vn_h, h + spread
