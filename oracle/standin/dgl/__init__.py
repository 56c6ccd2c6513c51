"""Stand-in for the subset of ``dgl==0.4.2`` that the reference's DGN path touches.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference's hot path calls into
DGL 0.4.2 (pinned in realworld_benchmark/environment_cpu.yml:14 and requirements.txt:8),
which is a third-party dependency whose source is NOT under /root/reference and which is
not installable here.  This module restates, from the published behaviour of that
release, exactly what these call sites rely on:

* realworld_benchmark/nets/dgn_layer.py:112,183,261   ``g.apply_edges(udf)``
* realworld_benchmark/nets/dgn_layer.py:115,186,264   ``g.update_all(message_udf, reduce_udf)``
* realworld_benchmark/nets/*/dgn_net.py               ``dgl.{sum,mean,max}_nodes(g, key)``
* realworld_benchmark/nets/dgn_layer.py:9,31,45       ``dgl.nn.pytorch.glob`` readouts,
                                                       ``g.batch_num_nodes`` as a list
* realworld_benchmark/data/molecules.py:229           ``dgl.batch(graphs)``

DGL-0.4.2 semantics restated (unverifiable here - flagged in DESIGN.md):
 (i)   ``edges.src[k] = ndata[k][src]``, ``edges.dst[k] = ndata[k][dst]``, ``edges.data`` = edata.
 (ii)  ``update_all`` uses *degree bucketing*: the reduce UDF runs once per distinct
       in-degree ``d > 0`` with ``nodes.mailbox[k]`` of shape ``[n_d, d, ...]``; the ``d``
       messages of one node are ordered by edge id; ``nodes.data`` holds those nodes' rows.
 (iii) the returned fields overwrite ``ndata`` for all nodes; nodes with in-degree 0 get
       zero rows.
 (iv)  ``batch_num_nodes`` / ``batch_num_edges`` are plain list attributes.
"""
from __future__ import annotations

import numpy as np
import torch

from . import nn  # noqa: F401  (makes ``dgl.nn.pytorch.glob`` importable)

__version__ = "0.4.2-standin"


class _EdgeView:
    def __init__(self, src, dst, data):
        self.src, self.dst, self.data = src, dst, data


class _NodeView:
    def __init__(self, data, mailbox):
        self.data, self.mailbox = data, mailbox


class _Gathered(dict):
    """Lazy ``{key: ndata[key][index]}`` so only the fields a UDF reads are gathered."""

    def __init__(self, frame, index):
        super().__init__()
        self._frame, self._index = frame, index

    def __missing__(self, key):
        val = self._frame[key].index_select(0, self._index.to(self._frame[key].device))
        self[key] = val
        return val


class DGLGraph:
    def __init__(self, num_nodes=0, src=None, dst=None, batch_num_nodes=None, batch_num_edges=None):
        self._n = int(num_nodes)
        self._src = torch.as_tensor(np.asarray(src if src is not None else []), dtype=torch.int64)
        self._dst = torch.as_tensor(np.asarray(dst if dst is not None else []), dtype=torch.int64)
        self.ndata, self.edata = {}, {}
        self.batch_num_nodes = list(batch_num_nodes) if batch_num_nodes is not None else [self._n]
        self.batch_num_edges = (list(batch_num_edges) if batch_num_edges is not None
                                else [int(self._src.numel())])
        self._buckets = None

    # --- structure queries -------------------------------------------------------------
    def number_of_nodes(self):
        return self._n

    def number_of_edges(self):
        return int(self._src.numel())

    def edges(self):
        return self._src, self._dst

    def in_degrees(self):
        return torch.bincount(self._dst, minlength=self._n)

    @property
    def batch_size(self):
        return len(self.batch_num_nodes)

    # --- message passing ---------------------------------------------------------------
    def apply_edges(self, func):
        ev = _EdgeView(_Gathered(self.ndata, self._src), _Gathered(self.ndata, self._dst), self.edata)
        self.edata.update(func(ev))

    def _degree_buckets(self):
        if self._buckets is None:
            deg = self.in_degrees()
            # stable sort by destination keeps edge-id order inside every mailbox row
            order = torch.sort(self._dst, stable=True)[1]
            start = torch.cumsum(deg, 0) - deg
            buckets = []
            for d in torch.unique(deg).tolist():
                if d == 0:
                    continue
                nodes = torch.nonzero(deg == d, as_tuple=False).flatten()
                slots = (start[nodes].unsqueeze(1) + torch.arange(d).unsqueeze(0)).flatten()
                buckets.append((d, nodes, order[slots]))
            self._buckets = buckets
        return self._buckets

    def update_all(self, message_func, reduce_func):
        ev = _EdgeView(_Gathered(self.ndata, self._src), _Gathered(self.ndata, self._dst), self.edata)
        msgs = message_func(ev)
        pieces, fields = [], None
        for d, nodes, eids in self._degree_buckets():
            box = {k: v.index_select(0, eids.to(v.device)).reshape((nodes.numel(), d) + tuple(v.shape[1:]))
                   for k, v in msgs.items()}
            out = reduce_func(_NodeView(_Gathered(self.ndata, nodes), box))
            fields = fields or list(out.keys())
            pieces.append((nodes, out))
        for k in (fields or []):
            proto = pieces[0][1][k]
            full = proto.new_zeros((self._n,) + tuple(proto.shape[1:]))
            idx = torch.cat([nodes for nodes, _ in pieces]).to(proto.device)
            full = full.index_copy(0, idx, torch.cat([out[k] for _, out in pieces], 0))
            self.ndata[k] = full


BatchedDGLGraph = DGLGraph


def batch(graphs):
    off, srcs, dsts = 0, [], []
    for g in graphs:
        s, d = g.edges()
        srcs.append(s + off)
        dsts.append(d + off)
        off += g.number_of_nodes()
    out = DGLGraph(off, torch.cat(srcs), torch.cat(dsts),
                   [g.number_of_nodes() for g in graphs], [g.number_of_edges() for g in graphs])
    for k in graphs[0].ndata:
        out.ndata[k] = torch.cat([g.ndata[k] for g in graphs], 0)
    for k in graphs[0].edata:
        out.edata[k] = torch.cat([g.edata[k] for g in graphs], 0)
    return out


def _segments(g, key):
    return torch.split(g.ndata[key], list(g.batch_num_nodes), dim=0)


def sum_nodes(g, key):
    return torch.stack([s.sum(0) for s in _segments(g, key)], 0)


def mean_nodes(g, key):
    return torch.stack([s.mean(0) for s in _segments(g, key)], 0)


def max_nodes(g, key):
    return torch.stack([s.max(0)[0] for s in _segments(g, key)], 0)
