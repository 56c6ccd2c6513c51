"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def samples_from_golden(gold):
    """Rebuild the per-graph sample dicts of a golden layer/net case."""
    sizes = gold["batch_num_nodes"].tolist()
    src, dst = gold["src"].astype(np.int64), gold["dst"].astype(np.int64)
    offs = np.concatenate([[0], np.cumsum(sizes)])
    out = []
    for gi, n in enumerate(sizes):
        lo, hi = offs[gi], offs[gi + 1]
        m = (dst >= lo) & (dst < hi)
        s = {"n": int(n), "src": (src[m] - lo).astype(np.int32), "dst": (dst[m] - lo).astype(np.int32),
             "eig": gold["eig"][lo:hi], "label": np.float32(0.0)}
        s["node_feat"] = gold["node_feat"][lo:hi] if "node_feat" in gold else np.zeros(n, np.int64)
        s["edge_feat"] = gold["edge_feat"][m] if "edge_feat" in gold else np.zeros(int(m.sum()), np.int64)
        out.append(s)
    return out


def state_from_golden(gold, fresh_module):
    """Parameters from the fixture; BatchNorm buffers keep the fresh module's (pre-step) values."""
    sd = fresh_module.state_dict()
    params = {k for k, _ in fresh_module.named_parameters()}
    for k in sd:
        if k in params:
            sd[k] = torch.from_numpy(gold["sd/" + k]).to(sd[k].device)
    return sd


def tol(ref, rel=1e-5):
    """north_star tolerance: |a-b| <= rel * max(1, ||ref||_inf)."""
    return rel * max(1.0, float(np.max(np.abs(ref))) if np.size(ref) else 1.0)


def assert_close(got, ref, rel=1e-5, what=""):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    ref = ref.detach().cpu().numpy() if isinstance(ref, torch.Tensor) else np.asarray(ref)
    assert got.shape == ref.shape, "%s shape %s vs %s" % (what, got.shape, ref.shape)
    err = float(np.max(np.abs(got.astype(np.float64) - ref.astype(np.float64)))) if got.size else 0.0
    assert err <= tol(ref, rel), "%s: max abs err %.3e > %.3e" % (what, err, tol(ref, rel))
