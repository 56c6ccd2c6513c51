"""``SCALERS`` registry with the reference's names (realworld_benchmark/nets/scalers.py:7-21).

Inside a layer the scalers are fused into the aggregation kernel's epilogue (one coefficient
per node and scaler, ``log(D+1)/avg`` or ``avg/log(D+1)``).  The registry values stay callable
with the reference signature ``scale(h, D, avg_d)`` for code that applies them by hand.
"""
from __future__ import annotations

import numpy as np

from dgn_b200 import _lib


class Scaler:
    def __init__(self, name, kind):
        self.name, self.kind = name, int(kind)

    def __repr__(self):
        return "Scaler(%s)" % self.name

    def __call__(self, h, D=None, avg_d=None):
        if self.kind == _lib.SCALE_IDENTITY:                     # scalers.py:7-8
            return h
        if self.kind == _lib.SCALE_AMPLIFICATION:                # scalers.py:11-13
            return h * (np.log(D + 1) / avg_d["log"])
        return h * (avg_d["log"] / np.log(D + 1))                # scalers.py:16-18


SCALERS = {"identity": Scaler("identity", _lib.SCALE_IDENTITY),
           "amplification": Scaler("amplification", _lib.SCALE_AMPLIFICATION),
           "attenuation": Scaler("attenuation", _lib.SCALE_ATTENUATION)}
