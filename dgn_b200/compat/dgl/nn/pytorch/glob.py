"""``dgl.nn.pytorch.glob`` names imported by realworld_benchmark/nets/dgn_layer.py:9."""
import dgl


def mean_nodes(g, key):
    return dgl.mean_nodes(g, key)


def sum_nodes(g, key):
    return dgl.sum_nodes(g, key)


def max_nodes(g, key):
    return dgl.max_nodes(g, key)
