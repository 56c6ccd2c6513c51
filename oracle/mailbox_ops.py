"""Aggregators and degree scalers of the reference, restated on a degree-bucketed mailbox.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function takes the mailbox of one
degree bucket: ``msg [n, D, F]``, the eigenvector rows of the source and destination of each
message ``eig_s, eig_d [n, D, K]`` and the destination's own features ``h_in [n, F]``, and
returns ``[n, F]`` - the calling convention of realworld_benchmark/nets/aggregators.py:8-71.
The op sequence of each formula is kept identical to the reference so results agree to the
last bit on the same torch build; only the code organisation differs (one shared
"directional weight" helper + a generated registry instead of 24 hand-written partials).
"""
from __future__ import annotations

import math

import numpy as np
import torch

EPS = 1e-8                      # realworld_benchmark/nets/aggregators.py:5


# ---------------------------------------------------------------------------------------------
# isotropic aggregators - aggregators.py:8-32
# ---------------------------------------------------------------------------------------------
def agg_mean(msg, eig_s, eig_d, h_in):          # :8-9
    return msg.mean(dim=1)


def agg_sum(msg, eig_s, eig_d, h_in):           # :31-32
    return msg.sum(dim=1)


def agg_max(msg, eig_s, eig_d, h_in):           # :12-13
    return msg.max(dim=1)[0]


def agg_min(msg, eig_s, eig_d, h_in):           # :16-17
    return msg.min(dim=1)[0]


def agg_var(msg, eig_s, eig_d, h_in):           # :24-28  relu(E[m^2] - E[m]^2)
    second = (msg * msg).mean(dim=-2)
    first = msg.mean(dim=-2)
    return torch.relu(second - first * first)


def agg_std(msg, eig_s, eig_d, h_in):           # :20-21
    return torch.sqrt(agg_var(msg, eig_s, eig_d, h_in) + EPS)


# ---------------------------------------------------------------------------------------------
# directional aggregators - aggregators.py:35-71
# ---------------------------------------------------------------------------------------------
def _field(eig_s, eig_d, k):
    """delta_uv = eig[u, k] - eig[v, k] for every message, shape [n, D]."""
    return eig_s[:, :, k] - eig_d[:, :, k]


def _l1_share(x):
    """x / (sum_D |x| + EPS): the normalisation used by av / dx / balanced."""
    return x / (torch.sum(torch.abs(x), keepdim=True, dim=1) + EPS)


def agg_dir_av(msg, eig_s, eig_d, h_in, k):                     # :35-39
    w = _l1_share(torch.abs(_field(eig_s, eig_d, k)))
    return torch.sum(torch.mul(msg, w.unsqueeze(-1)), dim=1)


def agg_dir_softmax(msg, eig_s, eig_d, h_in, k, alpha):         # :42-45
    w = torch.softmax(alpha * torch.abs(_field(eig_s, eig_d, k)).unsqueeze(-1), dim=1)
    return torch.sum(torch.mul(msg, w), dim=1)


def _derivative(msg, w, h_in):
    """sum_u w_u m_u - (sum_u w_u) h_in : the directional-derivative form (:51-52, :58-59, :70-71)."""
    w = w.unsqueeze(-1)
    return torch.sum(torch.mul(msg, w), dim=1) - torch.sum(w, dim=1) * h_in


def agg_dir_dx(msg, eig_s, eig_d, h_in, k):                     # :48-52
    return torch.abs(_derivative(msg, _l1_share(_field(eig_s, eig_d, k)), h_in))


def agg_dir_dx_no_abs(msg, eig_s, eig_d, h_in, k):              # :55-59
    return _derivative(msg, _l1_share(_field(eig_s, eig_d, k)), h_in)


def agg_dir_dx_balanced(msg, eig_s, eig_d, h_in, k):            # :62-71
    fwd = _l1_share(torch.relu(_field(eig_s, eig_d, k)))
    bwd = _l1_share(torch.relu(_field(eig_d, eig_s, k)))
    return torch.abs(_derivative(msg, (fwd + bwd) / 2, h_in))


def _bind(fn, **kw):
    def bound(msg, eig_s, eig_d, h_in):
        return fn(msg, eig_s, eig_d, h_in, **kw)
    bound.__name__ = fn.__name__ + "".join("_%s" % v for v in kw.values())
    return bound


def build_aggregator_registry(max_eig_idx: int = 3) -> dict:
    """Name -> callable, the 24 keys of aggregators.py:74-93 for ``max_eig_idx=3``."""
    reg = {"mean": agg_mean, "sum": agg_sum, "max": agg_max, "min": agg_min, "std": agg_std, "var": agg_var}
    for k in range(1, max_eig_idx + 1):
        reg["dir%d-av" % k] = _bind(agg_dir_av, k=k)
        reg["dir%d-0.1" % k] = _bind(agg_dir_softmax, k=k, alpha=0.1)
        reg["dir%d-neg-0.1" % k] = _bind(agg_dir_softmax, k=k, alpha=-0.1)
        reg["dir%d-dx" % k] = _bind(agg_dir_dx, k=k)
        reg["dir%d-dx-no-abs" % k] = _bind(agg_dir_dx_no_abs, k=k)
        reg["dir%d-dx-balanced" % k] = _bind(agg_dir_dx_balanced, k=k)
    return reg


AGGREGATORS = build_aggregator_registry(3)


# ---------------------------------------------------------------------------------------------
# degree scalers - realworld_benchmark/nets/scalers.py:7-21
# ---------------------------------------------------------------------------------------------
def scale_identity(h, D=None, avg_d=None):                      # :7-8
    return h


def scale_amplification(h, D, avg_d):                           # :11-13  h * log(D+1)/avg
    return h * (np.log(D + 1) / avg_d["log"])


def scale_attenuation(h, D, avg_d):                             # :16-18  h * avg/log(D+1)
    return h * (avg_d["log"] / np.log(D + 1))


SCALERS = {"identity": scale_identity, "amplification": scale_amplification, "attenuation": scale_attenuation}


def reduce_bucket(msg, eig_s, eig_d, h_in, aggregators, scalers, avg_d):
    """One degree bucket of ``reduce_func`` (realworld_benchmark/nets/dgn_layer.py:86-98).

    Output layout ``[n, S*A*F]``: scaler-major, then aggregator, then feature.  A single
    scaler is *not applied* (dgn_layer.py:95 only scales when ``len(scalers) > 1``).
    """
    D = msg.shape[-2]
    out = torch.cat([agg(msg, eig_s, eig_d, h_in) for agg in aggregators], dim=1)
    if len(scalers) > 1:
        out = torch.cat([sc(out, D=D, avg_d=avg_d) for sc in scalers], dim=1)
    return out


def log_degree_factor(D: int) -> float:
    return math.log(D + 1)
