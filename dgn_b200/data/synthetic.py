"""Seeded synthetic mini-batches shaped like the reference's datasets.

There are no datasets (and no network) in the build or GPU containers, so every
measurement and parity test runs on graphs generated here.  The generators follow
SURVEY.md section 8(d); the statistics they imitate come from the reference's
loaders:

* ZINC-like      - realworld_benchmark/data/molecules.py:58-116 (28 atom types, 3 bond
                   types, 6 stored eigenvector columns, ``L = diag(clip(deg,1)) - A``)
* CIFAR10-like   - realworld_benchmark/data/superpixels.py:50-69,423-428 (directed 8-NN,
                   ``eig = [0, x, y]``)
* PATTERN-like   - realworld_benchmark/data/SBMs.py:110-139,158 (5 stored columns)
* molhiv-like    - realworld_benchmark/data/HIV.py:17-46,56,66 (4 stored columns)

Everything is plain numpy on the host; a graph sample is a small dict of arrays.  Edges
are listed in edge-id order (sorted by source, then destination), which is the order the
reference's mailbox sees them in.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "laplacian_eigvecs",
    "zinc_like_graph",
    "cifar_like_graph",
    "pattern_like_graph",
    "molhiv_like_graph",
    "make_samples",
    "avg_log_degree",
]


def _edges_from_adj(adj: np.ndarray):
    """Directed edge list (src, dst) of a dense 0/1 matrix, ordered by (src, dst)."""
    src, dst = np.nonzero(adj)
    return src.astype(np.int32), dst.astype(np.int32)


def laplacian_eigvecs(n: int, src: np.ndarray, dst: np.ndarray, k: int, rng=None) -> np.ndarray:
    """First ``k`` eigenvectors (ascending eigenvalue) of ``L = diag(clip(in_deg,1)) - A``.

    Mirrors the *intent* of realworld_benchmark/data/molecules.py:100-116 (norm='none').
    The reference uses ARPACK with ``tol=5e-1`` whose output is not reproducible; a dense
    ``eigh`` on the symmetrised Laplacian gives the same subspace deterministically.
    Columns get a random sign when ``rng`` is given (ARPACK's sign is arbitrary too).
    Graphs with fewer than ``k`` nodes are zero padded on the right.
    """
    a = np.zeros((n, n), dtype=np.float64)
    a[src, dst] = 1.0
    a = np.maximum(a, a.T)
    deg = np.clip(a.sum(0), 1.0, None)
    lap = np.diag(deg) - a
    _, vec = np.linalg.eigh(lap)
    out = np.zeros((n, k), dtype=np.float32)
    kk = min(k, n)
    out[:, :kk] = vec[:, :kk].astype(np.float32)
    if rng is not None:
        out *= rng.choice(np.array([-1.0, 1.0], dtype=np.float32), size=(1, k))
    return out


def _molecule_adj(n: int, rng, target_ratio: float = 2.15, max_deg: int = 4) -> np.ndarray:
    """Connected, symmetric, max-degree-limited adjacency with ~target_ratio directed edges/node."""
    adj = np.zeros((n, n), dtype=np.uint8)
    deg = np.zeros(n, dtype=np.int64)
    order = rng.permutation(n)
    for i in range(1, n):
        v = order[i]
        cand = [u for u in order[:i] if deg[u] < max_deg]
        u = cand[rng.integers(len(cand))] if cand else order[rng.integers(i)]
        adj[u, v] = adj[v, u] = 1
        deg[u] += 1
        deg[v] += 1
    want = int(round(target_ratio * n / 2.0))
    have = n - 1
    tries = 0
    while have < want and tries < 20 * n:
        tries += 1
        u, v = rng.integers(n), rng.integers(n)
        if u == v or adj[u, v] or deg[u] >= max_deg or deg[v] >= max_deg:
            continue
        adj[u, v] = adj[v, u] = 1
        deg[u] += 1
        deg[v] += 1
        have += 1
    return adj


def zinc_like_graph(rng, k_eig: int = 6, n_min: int = 9, n_max: int = 37) -> dict:
    n = int(rng.integers(n_min, n_max + 1))
    src, dst = _edges_from_adj(_molecule_adj(n, rng))
    return {
        "n": n,
        "src": src,
        "dst": dst,
        "node_feat": rng.integers(0, 28, size=n).astype(np.int64),
        "edge_feat": rng.integers(1, 4, size=src.shape[0]).astype(np.int64),
        "eig": laplacian_eigvecs(n, src, dst, k_eig, rng),
        "label": np.float32(rng.standard_normal()),
    }


def molhiv_like_graph(rng, k_eig: int = 4) -> dict:
    n = int(np.clip(np.round(rng.lognormal(mean=3.15, sigma=0.42)), 6, 222))
    src, dst = _edges_from_adj(_molecule_adj(n, rng, target_ratio=2.15, max_deg=6))
    node_dims = np.array([119, 4, 12, 12, 10, 6, 6, 2, 2])
    edge_dims = np.array([5, 6, 2])
    return {
        "n": n,
        "src": src,
        "dst": dst,
        "node_feat": (rng.random((n, 9)) * node_dims).astype(np.int64),
        "edge_feat": (rng.random((src.shape[0], 3)) * edge_dims).astype(np.int64),
        "eig": laplacian_eigvecs(n, src, dst, k_eig, rng),
        "label": np.float32(rng.integers(0, 2)),
    }


def cifar_like_graph(rng, knn: int = 8, n_min: int = 85, n_max: int = 150) -> dict:
    n = int(rng.integers(n_min, n_max + 1))
    xy = rng.random((n, 2)).astype(np.float32)
    d2 = ((xy[:, None, :] - xy[None, :, :]) ** 2).sum(-1)
    np.fill_diagonal(d2, np.inf)
    nbr = np.argsort(d2, axis=1, kind="stable")[:, :knn]
    adj = np.zeros((n, n), dtype=np.uint8)
    adj[np.repeat(np.arange(n), knn), nbr.reshape(-1)] = 1   # node -> its k nearest (directed)
    src, dst = _edges_from_adj(adj)
    eig = np.concatenate([np.zeros((n, 1), np.float32), xy], axis=1)
    return {
        "n": n,
        "src": src,
        "dst": dst,
        "node_feat": rng.random((n, 5)).astype(np.float32),
        "edge_feat": np.sqrt(d2[src, dst]).astype(np.float32)[:, None],
        "eig": eig,
        "label": np.int64(rng.integers(0, 10)),
    }


def pattern_like_graph(rng, k_eig: int = 5, n_min: int = 100, n_max: int = 180,
                       p: float = 0.5, q: float = 0.35) -> dict:
    n = int(rng.integers(n_min, n_max + 1))
    block = rng.integers(0, 5, size=n)
    prob = np.where(block[:, None] == block[None, :], p, q)
    upper = np.triu(rng.random((n, n)) < prob, 1)
    adj = (upper | upper.T).astype(np.uint8)
    lonely = np.nonzero(adj.sum(0) == 0)[0]
    for v in lonely:                                   # keep every node reachable
        u = (v + 1) % n
        adj[u, v] = adj[v, u] = 1
    src, dst = _edges_from_adj(adj)
    return {
        "n": n,
        "src": src,
        "dst": dst,
        "node_feat": rng.integers(0, 3, size=n).astype(np.int64),
        "edge_feat": np.ones((src.shape[0], 1), np.float32),
        "eig": laplacian_eigvecs(n, src, dst, k_eig, rng),
        "label": (block == 0).astype(np.int64),      # per-node labels (node classification)
    }


_KINDS = {
    "zinc": zinc_like_graph,
    "molhiv": molhiv_like_graph,
    "cifar": cifar_like_graph,
    "pattern": pattern_like_graph,
}


def make_samples(kind: str, n_graphs: int, seed: int = 0, **kw) -> list:
    """``n_graphs`` independent samples of ``kind`` from ``numpy.random.default_rng(seed)``."""
    rng = np.random.default_rng(seed)
    fn = _KINDS[kind]
    return [fn(rng, **kw) for _ in range(n_graphs)]


def avg_log_degree(samples) -> float:
    """``mean(log(in_degree + 1))`` over all nodes: the ``avg_d['log']`` statistic.

    Mirrors realworld_benchmark/main_molecules.py:300-304.
    """
    acc, cnt = 0.0, 0
    for s in samples:
        deg = np.bincount(s["dst"], minlength=s["n"]).astype(np.float32)
        acc += float(np.log(deg + 1.0).sum())
        cnt += s["n"]
    return acc / max(cnt, 1)
