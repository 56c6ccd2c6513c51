"""DGN layers with the reference's constructors, parameter names and forward signatures,
executed by the fused sm_100a aggregation kernels.

Drop-in for realworld_benchmark/nets/dgn_layer.py: ``DGNLayer(...).model.forward(g, h, e, snorm_n)``
(:328-352, :103, :178, :309), sub-module names ``pretrans`` / ``posttrans`` / ``batchnorm_h`` /
``towers`` / ``mixing_network`` (:66-70, :300-307) so ``state_dict``s interchange.

What changes is *how* a layer runs:

* no ``apply_edges`` / ``update_all`` / degree bucketing: one ``dgn_agg_forward`` launch walks the
  batched CSR and produces every aggregator x scaler slab (and the ``cat([h, agg])`` copy);
* a 1-layer ``pretrans`` is affine, so the edge-level GEMM over ``cat(h_u, h_v)`` is replaced by two
  node-level GEMMs ``P = h W_src^T``, ``Q = h W_dst^T + b`` and the kernel forms ``P[u] + Q[v]`` on the
  fly (``DGN_MSG_AFFINE``); deeper ``pretrans`` MLPs materialise ``[E, F]`` messages (``DGN_MSG_DENSE``);
* ``* snorm_n -> BatchNorm1d -> ReLU -> + h_in`` is one fused op (``dgn_norm_forward``).

The GEMMs stay plain fp32 library calls (TF32 would break the 1e-5 parity bound).
"""
from __future__ import annotations

import weakref

import torch
import torch.nn as nn
import torch.nn.functional as F

from dgn_b200 import _lib
from dgn_b200.fused import fused_layer
from dgn_b200.ops import AggSpec, PostSpec, aggregate, norm_act, readout

from .aggregators import AGGREGATORS
from .layers import MLP, FCLayer
from .scalers import SCALERS

EPS = 1e-5      # rb/nets/dgn_layer.py:1 (unused there as well; the live EPS is aggregators.EPS)

# producer layer -> weak reference to the pretrans Linear of the layer that consumed its output last time (see
# _FusedConv._fused_forward); kept outside the modules so that state_dict / deepcopy / pickle never see it
_NEXT_PRE = weakref.WeakKeyDictionary()


def _avg_log(avg_d) -> float:
    v = avg_d["log"]
    return float(v.item() if isinstance(v, torch.Tensor) else v)


class _FusedConv(nn.Module):
    """Shared engine-side plumbing of the simple / complex / tower layers."""

    def _init_common(self, in_dim, dropout, graph_norm, batch_norm, aggregators, scalers, avg_d):
        self.in_dim = in_dim
        self.dropout, self.graph_norm, self.batch_norm = dropout, graph_norm, batch_norm
        self.aggregators, self.scalers, self.avg_d = aggregators, scalers, avg_d
        self._specs = {}

    def _spec(self, n_eig) -> AggSpec:
        sp = self._specs.get(n_eig)
        if sp is None:
            sp = AggSpec(self.aggregators, self.scalers, _avg_log(self.avg_d), self.in_dim, n_eig)
            self._specs[n_eig] = sp
        return sp

    def _folded(self, n_eig, lead, n_out):
        """(raw-aggregate spec, PostSpec) of the scaler-folded path, or (None, None) when the shapes are outside it."""
        key = ("fold", n_eig, lead, n_out)
        ent = self._specs.get(key)
        if ent is None:
            spec = self._spec(n_eig)
            post = PostSpec(spec, lead, n_out)
            if post.supported(spec):
                raw = AggSpec(self.aggregators, [SCALERS["identity"]], spec.avg_log, self.in_dim, n_eig)
                ent = (raw, post)
            else:
                ent = (None, None)
            self._specs[key] = ent
        return ent

    @staticmethod
    def _eig(g, like):
        eig = g.ndata["eig"]
        return eig if eig.device == like.device else eig.to(like.device)

    @staticmethod
    def _affine(mlp):
        """The single Linear of a 1-layer MLP without activation / dropout / batch-norm, else None."""
        fcs = mlp.fully_connected
        if len(fcs) == 1 and fcs[0].activation is None and fcs[0].dropout is None and fcs[0].b_norm is None \
                and fcs[0].linear.bias is not None:
            return fcs[0].linear
        return None

    def _fused_forward(self, g, h, e, snorm_n, relu, residual):
        """Whole layer as one autograd node (dgn_b200/fused.py) when pre/posttrans are single Linears."""
        post = self._affine(self.posttrans)
        pre = self._affine(self.pretrans) if hasattr(self, "pretrans") else None
        if post is None or (hasattr(self, "pretrans") and pre is None) or not h.is_cuda:
            return None
        if pre is not None and (self.in_dim % 4 or post.out_features % 4) and not getattr(self, "_in_padded", False):
            # widths off the 16 B grid (the reference's 45 / 47 / 65 ...): layer-owned zero-padded operands
            pad = self.__dict__.get("_padded")
            if pad is None:
                from dgn_b200.towers import TowerFusion
                pad = self.__dict__["_padded"] = TowerFusion([self], self.in_dim, post.out_features, relu=relu,
                                                             residual=residual)
            if pad.supported(h):
                return pad.forward(g, h, snorm_n)
        eig = self._eig(g, h)
        R = None
        if pre is not None and self.edge_features:
            R = F.linear(e, pre.weight[:, 2 * self.in_dim:])            # per-edge term W_e ef, edge-id order
        spec_raw, pspec = self._folded(eig.shape[1], self.in_dim if pre is not None else 0, post.out_features)
        if pspec is not None and post.in_features != pspec.w_cols:
            spec_raw = pspec = None
        # Cross-layer fusion without touching the caller's loop (``for conv in self.layers: h = conv(g, h, e, snorm_n)``,
        # rb/nets/*/dgn_net.py): a layer tags its output with itself; the layer that receives the tensor registers its
        # pretrans weight with the producer, whose epilogue from then on also computes this layer's P = h W_src^T,
        # Q = h W_dst^T (one launch less per layer) and hands them over on the tensor.  Anything that does not match
        # (another consumer, changed weights, different graph) simply recomputes.
        pq = None
        if pre is not None:
            tag = getattr(h, "_dgn_pq", None)
            if tag is not None and tag[0] is pre.weight and tag[1] == pre.weight._version and tag[2] is g:
                pq = tag[3]
            producer = getattr(h, "_dgn_from", None)
            if producer is not None and producer() is not None:
                _NEXT_PRE[producer()] = weakref.ref(pre)
        ref = _NEXT_PRE.get(self)
        nxt = ref() if ref is not None else None
        next_w = None
        if nxt is not None and nxt.weight.is_cuda and nxt.weight.dtype == torch.float32 and not (self.dropout and self.training):
            next_w = nxt.weight
        out, pq_out = fused_layer(g, self._spec(eig.shape[1]), eig, h, R, pre, post,
                                  self.batchnorm_h if self.batch_norm else None, snorm_n if self.graph_norm else None,
                                  self.training, relu, residual, self.in_dim, spec_raw=spec_raw, post=pspec, pq=pq,
                                  next_w=next_w)
        if self.dropout and self.training:
            out = F.dropout(out, self.dropout, training=True)
        out._dgn_from = weakref.ref(self)
        if pq_out is not None:
            out._dgn_pq = (next_w, next_w._version, g, pq_out)
        return out

    def _pretrans_aggregate(self, g, h, e, cat_input=True):
        """messages = pretrans(cat(h_u, h_v[, ef])) (rb/nets/dgn_layer.py:75-80), then all aggregators."""
        eig = self._eig(g, h)
        spec = self._spec(eig.shape[1])
        fcs = self.pretrans.fully_connected
        Fi = self.in_dim
        affine = (len(fcs) == 1 and fcs[0].activation is None and fcs[0].dropout is None and fcs[0].b_norm is None)
        if affine:
            lin = fcs[0].linear
            W = lin.weight                                        # [F, 2F (+edge_dim)]
            P = F.linear(h, W[:, :Fi])                            # source half, gathered per edge in-kernel
            Q = F.linear(h, W[:, Fi:2 * Fi], lin.bias)            # destination half (+ bias), once per node
            R = F.linear(e, W[:, 2 * Fi:]) if self.edge_features else None
            return aggregate(g, spec, _lib.MSG_AFFINE, h, eig, x=P, q=Q, r=R, cat_input=cat_input)
        src, dst = g.edges()
        parts = [h.index_select(0, src), h.index_select(0, dst)] + ([e] if self.edge_features else [])
        M = self.pretrans(torch.cat(parts, dim=1))                # [E, F] in edge-id order
        return aggregate(g, spec, _lib.MSG_DENSE, h, eig, r=M, cat_input=cat_input)

    def _epilogue(self, g, y, snorm_n, relu, residual):
        out = norm_act(y, snorm_n if self.graph_norm else None, self.batchnorm_h if self.batch_norm else None,
                       self.training, relu, residual, getattr(g, "n_rows_dev", None))
        if self.dropout and self.training:
            out = F.dropout(out, self.dropout, training=True)
        return out


class DGNLayerComplex(_FusedConv):
    def __init__(self, in_dim, out_dim, dropout, graph_norm, batch_norm, aggregators, scalers, avg_d, residual,
                 edge_features, edge_dim, pretrans_layers=1, posttrans_layers=1):
        super().__init__()
        self._init_common(in_dim, dropout, graph_norm, batch_norm, aggregators, scalers, avg_d)
        self.edge_features = bool(edge_features)
        self.residual = residual
        self.batchnorm_h = nn.BatchNorm1d(out_dim)
        self.pretrans = MLP(in_size=2 * in_dim + (edge_dim if edge_features else 0), hidden_size=in_dim,
                            out_size=in_dim, layers=pretrans_layers, mid_activation="relu", last_activation="none")
        self.posttrans = MLP(in_size=(len(aggregators) * len(scalers) + 1) * in_dim, hidden_size=out_dim,
                             out_size=out_dim, layers=posttrans_layers, mid_activation="relu", last_activation="none")
        if in_dim != out_dim:
            self.residual = False

    def forward(self, g, h, e, snorm_n):
        out = self._fused_forward(g, h, e, snorm_n, relu=True, residual=bool(self.residual))
        if out is not None:
            return out
        y = self.posttrans(self._pretrans_aggregate(g, h, e, cat_input=True))
        return self._epilogue(g, y, snorm_n, relu=True, residual=h if self.residual else None)


class DGNLayerSimple(_FusedConv):
    def __init__(self, in_dim, out_dim, dropout, graph_norm, batch_norm, aggregators, scalers, residual, avg_d,
                 posttrans_layers=1):
        super().__init__()
        self._init_common(in_dim, dropout, graph_norm, batch_norm, aggregators, scalers, avg_d)
        self.residual = residual
        self.batchnorm_h = nn.BatchNorm1d(out_dim)
        self.posttrans = MLP(in_size=(len(aggregators) * len(scalers)) * in_dim, hidden_size=out_dim,
                             out_size=out_dim, layers=posttrans_layers, mid_activation="relu", last_activation="none")
        if in_dim != out_dim:
            self.residual = False

    def forward(self, g, h, e, snorm_n):
        out = self._fused_forward(g, h, e, snorm_n, relu=True, residual=bool(self.residual))
        if out is not None:
            return out
        eig = self._eig(g, h)
        agg = aggregate(g, self._spec(eig.shape[1]), _lib.MSG_SOURCE, h, eig, x=h)   # message = h[src]
        y = self.posttrans(agg)
        return self._epilogue(g, y, snorm_n, relu=True, residual=h if self.residual else None)


class DGNTower(_FusedConv):
    def __init__(self, in_dim, out_dim, dropout, graph_norm, batch_norm, aggregators, scalers, avg_d,
                 pretrans_layers, posttrans_layers, edge_features, edge_dim):
        super().__init__()
        self._init_common(in_dim, dropout, graph_norm, batch_norm, aggregators, scalers, avg_d)
        self.edge_features = bool(edge_features)
        self.batchnorm_h = nn.BatchNorm1d(out_dim)
        self.pretrans = MLP(in_size=2 * in_dim + (edge_dim if edge_features else 0), hidden_size=in_dim,
                            out_size=in_dim, layers=pretrans_layers, mid_activation="relu", last_activation="none")
        self.posttrans = MLP(in_size=(len(aggregators) * len(scalers) + 1) * in_dim, hidden_size=out_dim,
                             out_size=out_dim, layers=posttrans_layers, mid_activation="relu", last_activation="none")

    def forward(self, g, h, e, snorm_n):
        out = self._fused_forward(g, h, e, snorm_n, relu=False, residual=False)   # no ReLU / residual in a tower
        if out is not None:
            return out
        y = self.posttrans(self._pretrans_aggregate(g, h, e, cat_input=True))
        return self._epilogue(g, y, snorm_n, relu=False, residual=None)


class DGNLayerTower(nn.Module):
    def __init__(self, in_dim, out_dim, aggregators, scalers, avg_d, dropout, graph_norm, batch_norm, towers=5,
                 pretrans_layers=1, posttrans_layers=1, divide_input=True, residual=False, edge_features=False,
                 edge_dim=0):
        super().__init__()
        assert ((not divide_input) or in_dim % towers == 0), "if divide_input is set the number of towers has to divide in_dim"
        assert (out_dim % towers == 0), "the number of towers has to divide the out_dim"
        assert avg_d is not None
        self.divide_input = divide_input
        self.input_tower = in_dim // towers if divide_input else in_dim
        self.output_tower = out_dim // towers
        self.in_dim, self.out_dim = in_dim, out_dim
        self.edge_features = edge_features
        self.residual = residual and in_dim == out_dim
        self.towers = nn.ModuleList(
            DGNTower(in_dim=self.input_tower, out_dim=self.output_tower, aggregators=aggregators, scalers=scalers,
                     avg_d=avg_d, pretrans_layers=pretrans_layers, posttrans_layers=posttrans_layers,
                     batch_norm=batch_norm, dropout=dropout, graph_norm=graph_norm, edge_features=edge_features,
                     edge_dim=edge_dim) for _ in range(towers))
        self.mixing_network = FCLayer(out_dim, out_dim, activation="LeakyReLU")

    def forward(self, g, h, e, snorm_n):
        w = self.input_tower
        fusion = self.__dict__.get("_fusion")
        if fusion is None:
            from dgn_b200.towers import TowerFusion
            fusion = self.__dict__["_fusion"] = TowerFusion(self.towers, self.input_tower, self.output_tower)
        if len(self.towers) > 1 and self.divide_input and fusion.supported(h):
            # all towers as ONE block-structured layer: one aggregation launch, one posttrans GEMM, one epilogue
            y = fusion.forward(g, h, snorm_n)
        elif self.divide_input:
            y = torch.cat([tw(g, h[:, i * w:(i + 1) * w], e, snorm_n) for i, tw in enumerate(self.towers)], dim=1)
        else:
            y = torch.cat([tw(g, h, e, snorm_n) for tw in self.towers], dim=1)
        if len(self.towers) > 1:
            y = self.mixing_network(y)
        return h + y if self.residual else y


class DGNLayer(nn.Module):
    """Factory: resolves the registry names, builds ``self.model`` (callers use ``.model``)."""

    def __init__(self, in_dim, out_dim, dropout, graph_norm, batch_norm, aggregators, scalers, avg_d, type_net,
                 residual, towers=5, divide_input=True, edge_features=None, edge_dim=None, pretrans_layers=1,
                 posttrans_layers=1):
        super().__init__()
        aggregators = [AGGREGATORS[aggr] for aggr in aggregators.split()]     # unknown name -> KeyError
        scalers = [SCALERS[scale] for scale in scalers.split()]
        if type_net == "simple":
            self.model = DGNLayerSimple(in_dim=in_dim, out_dim=out_dim, dropout=dropout, graph_norm=graph_norm,
                                        batch_norm=batch_norm, residual=residual, aggregators=aggregators,
                                        scalers=scalers, avg_d=avg_d, posttrans_layers=posttrans_layers)
        elif type_net == "complex":
            self.model = DGNLayerComplex(in_dim=in_dim, out_dim=out_dim, dropout=dropout, graph_norm=graph_norm,
                                         batch_norm=batch_norm, aggregators=aggregators, residual=residual,
                                         scalers=scalers, avg_d=avg_d, edge_features=edge_features,
                                         edge_dim=edge_dim, pretrans_layers=pretrans_layers,
                                         posttrans_layers=posttrans_layers)
        elif type_net == "towers":
            self.model = DGNLayerTower(in_dim=in_dim, out_dim=out_dim, aggregators=aggregators, scalers=scalers,
                                       avg_d=avg_d, dropout=dropout, graph_norm=graph_norm, batch_norm=batch_norm,
                                       towers=towers, pretrans_layers=pretrans_layers,
                                       posttrans_layers=posttrans_layers, divide_input=divide_input,
                                       residual=residual, edge_features=edge_features, edge_dim=edge_dim)


class VirtualNode(nn.Module):
    """Per-graph pooled "virtual node" (rb/nets/dgn_layer.py:12-49; used by the PCBA net only)."""

    def __init__(self, dim, dropout, batch_norm=False, bias=True, residual=True, vn_type="mean"):
        super().__init__()
        self.vn_type = vn_type.lower()
        self.fc_layer = FCLayer(in_size=dim, out_size=dim, activation="relu", dropout=dropout, b_norm=batch_norm,
                                bias=bias)
        self.residual = residual

    def forward(self, g, h, vn_h):
        if self.vn_type == "mean":
            pool = readout(g, h, "mean")
        elif self.vn_type == "sum":
            pool = readout(g, h, "sum")
        elif self.vn_type == "logsum":
            lognum = torch.log(torch.tensor(g.batch_num_nodes, dtype=h.dtype, device=h.device))
            pool = readout(g, h, "mean") * lognum.unsqueeze(-1)
        else:
            raise ValueError('Undefined input "%s". Accepted values are "sum", "mean", "logsum"' % self.vn_type)
        vn_new = self.fc_layer(vn_h + pool)
        vn_h = vn_h + vn_new if self.residual else vn_new
        counts = torch.as_tensor(g.batch_num_nodes, device=h.device)
        return vn_h, h + torch.repeat_interleave(vn_h, counts, dim=0)
