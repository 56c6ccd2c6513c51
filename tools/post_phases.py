#!/usr/bin/env python
"""Phase timeline inside dgn::post_fwd_kernel (per-CTA %globaltimer stamps, DGN_POST_DBG): where a ~20 us launch goes."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
dbg = torch.zeros(4096 * 8, dtype=torch.int64, device="cuda")
os.environ["DGN_POST_DBG"] = "%x" % dbg.data_ptr()

from dgn_b200 import ops  # noqa: E402
from dgn_b200.graph import BatchedGraph  # noqa: E402
from dgn_b200.nets.aggregators import AGGREGATORS  # noqa: E402
from dgn_b200.nets.scalers import SCALERS  # noqa: E402

N, F, A, Fo = 3136, 64, 10, 64
rng = np.random.default_rng(0)
deg = rng.integers(1, 5, size=N)
dst = np.repeat(np.arange(N), deg).astype(np.int32)
src = rng.integers(0, N, size=dst.shape[0]).astype(np.int32)
g = BatchedGraph(N, src, dst).to("cuda")
spec = ops.AggSpec([AGGREGATORS["mean"]] * A, [SCALERS[s] for s in ("identity", "amplification", "attenuation")], 1.2, F, 3)
ps = ops.PostSpec(spec, F, Fo)
cat = torch.randn(N, F + A * F, device="cuda")
W = torch.randn(Fo, ps.w_cols, device="cuda") / 8
y = torch.empty(N, Fo, device="cuda")
stats = torch.empty(640 * Fo, device="cuda")
bn = torch.nn.BatchNorm1d(Fo).cuda()
bias = torch.zeros(Fo, device="cuda")
for it in range(5):
    dbg.zero_()
    torch.cuda.synchronize()
    ops.post_forward(ps, g, cat, W, y, stats, bias, None, None)
    torch.cuda.synchronize()
t = dbg.cpu().numpy().reshape(-1, 8)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
names = ["start", "prologue done", "loader loop done", "accumulators ready", "after cluster sync 1", "reduce+stats done", "exit"]
print("CTAs: %d; kernel span %.2f us" % (len(t), (t[:, 6].max() - t0) / 1e3))
for i, nme in enumerate(names):
    col = (t[:, i] - t0) / 1e3
    print("%-24s mean %6.2f  min %6.2f  max %6.2f us" % (nme, col.mean(), col.min(), col.max()))
