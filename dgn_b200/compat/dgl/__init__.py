"""Minimal ``dgl`` surface for running the reference's task nets on ``dgn_b200.graph.BatchedGraph``
when real DGL is not installed (put ``dgn_b200/compat`` on ``sys.path``).

Only what realworld_benchmark/nets/*/dgn_net.py and nets/dgn_layer.py import: ``dgl.batch`` is NOT here
(collation is ``dgn_b200.graph.collate``); the readouts run the ``dgn_readout_*`` CUDA kernels.
"""
from dgn_b200.graph import BatchedGraph as DGLGraph  # noqa: F401
from dgn_b200.ops import readout as _readout

from . import nn  # noqa: F401

__version__ = "0.4.2-dgn_b200-compat"


def sum_nodes(g, key):
    return _readout(g, g.ndata[key], "sum")


def mean_nodes(g, key):
    return _readout(g, g.ndata[key], "mean")


def max_nodes(g, key):
    return _readout(g, g.ndata[key], "max")
