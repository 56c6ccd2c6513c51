#!/usr/bin/env python
"""Times dgn_gemm_tf32x3 against the fp32 library GEMM on the DGN layer shapes (CUDA graph of 20 calls)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgn_b200 import ops  # noqa: E402

CASES = [  # name, M, N, K, a_kmajor, b_kmajor, c_transposed
    ("y=cat@Wpost^T", 3008, 64, 1984, True, True, False),
    ("dcat=dy@Wpost", 3008, 1984, 64, True, False, False),
    ("dWpost=(cat^T@dy)^T", 1984, 64, 3008, False, False, True),
    ("P=h@Ws^T", 3008, 64, 64, True, True, False),
    ("dh+=dP@Ws", 3008, 64, 64, True, False, False),
    ("dWs=dP^T@h", 64, 64, 3008, False, False, False),
]


def timed(fn, reps=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (5 * reps)


def main():
    dev = "cuda"
    for name, M, N, K, ak, bk, ct in CASES:
        a = torch.randn((M, K) if ak else (K, M), device=dev)
        b = torch.randn((N, K) if bk else (K, N), device=dev)
        out = torch.empty((N, M) if ct else (M, N), device=dev)
        A = a if ak else a.t()
        B = b.t() if bk else b
        ref = torch.empty(M, N, device=dev)
        t_mine = timed(lambda: ops.gemm(a, b, a_kmajor=ak, b_kmajor=bk, out=out, c_transposed=ct))
        t_lib = timed(lambda: torch.mm(A, B, out=ref))
        flops = 2.0 * M * N * K
        print("%-22s M=%5d N=%5d K=%5d  tcgen05 %7.2f us (%6.1f TF/s eff)   library fp32 %7.2f us" %
              (name, M, N, K, t_mine, flops / t_mine / 1e6, t_lib))


if __name__ == "__main__":
    main()
