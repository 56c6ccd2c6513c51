"""Collate synthetic samples into a stand-in DGL batch (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows ``MoleculeDataset.collate`` (realworld_benchmark/data/molecules.py:219-230):
``dgl.batch`` of the graphs plus ``snorm_n = sqrt(1/n_g)`` per node, ``snorm_e = sqrt(1/e_g)`` per edge.
"""
from __future__ import annotations

import numpy as np
import torch

from . import use_standin_dgl


def collate_standin(samples):
    dgl = use_standin_dgl()
    graphs = []
    for s in samples:
        g = dgl.DGLGraph(s["n"], s["src"], s["dst"])
        g.ndata["feat"] = torch.from_numpy(np.asarray(s["node_feat"]))
        g.ndata["eig"] = torch.from_numpy(np.asarray(s["eig"]))
        g.edata["feat"] = torch.from_numpy(np.asarray(s["edge_feat"]))
        graphs.append(g)
    big = dgl.batch(graphs)
    snorm_n = torch.cat([torch.full((s["n"], 1), 1.0 / float(s["n"])) for s in samples]).sqrt()
    snorm_e = torch.cat([torch.full((len(s["src"]), 1), 1.0 / float(max(len(s["src"]), 1))) for s in samples]).sqrt()
    labels = np.asarray([s["label"] for s in samples]) if np.ndim(samples[0]["label"]) == 0 \
        else np.concatenate([s["label"] for s in samples])
    return big, torch.from_numpy(labels), snorm_n, snorm_e
