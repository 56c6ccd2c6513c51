#!/usr/bin/env python
"""Roofline sweep of the fused aggregation kernels over batch scale and BASELINE shapes (GPU only).

For every case the forward and backward launches are timed alone: a CUDA graph holds `rot` launches on
rotating operand sets (total footprint > 2x L2 where memory allows) and is replayed between CUDA events.
Prints one JSON line per case; `profiles/` keeps the table of the round.

    python tools/agg_sweep.py [--cases zinc,pattern,cifar,molhiv] [--scales 1,4,16,64]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from dgn_b200 import _lib, ops                                           # noqa: E402
from dgn_b200.data.synthetic import make_samples, avg_log_degree    # noqa: E402
from dgn_b200.graph import collate                                  # noqa: E402
from dgn_b200.nets.aggregators import AGGREGATORS                   # noqa: E402
from dgn_b200.nets.scalers import SCALERS                           # noqa: E402
from dgn_b200.ops import AggSpec, agg_backward_raw, agg_forward_raw  # noqa: E402

S3 = "identity amplification attenuation"
CASES = {
    # name: (generator kind, graphs at scale 1, F, aggregators, scalers, eig columns used)
    "zinc": ("zinc", 128, 64, "mean max min std dir1-dx dir2-dx dir1-dx-no-abs dir2-dx-no-abs dir1-av dir2-av", S3, 2),
    "cifar": ("cifar", 128, 64, "mean dir1-dx dir2-dx", "identity", 2),
    "molhiv": ("molhiv", 512, 80, "mean max min dir1-dx dir2-dx dir1-av dir2-av", "identity", 2),
    "pattern": ("pattern", 256, 48, "mean dir1-dx dir2-dx dir3-dx dir4-dx", S3, 4),
}


def peak():
    try:
        return float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def time_graph(fn, sets, replays=5):
    dev = sets[0]["h"].device
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        fn(sets[0])
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg):
        for t in sets:
            fn(t)
    for _ in range(2):
        cg.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(replays):
        cg.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / (replays * len(sets))


def run_case(name, scale, dev, folded=False):
    """``folded``: the layout the fused layer runs since round 2 - raw aggregates [N, F + A*F] (the scalers are folded
    into the posttrans GEMM, dgn_post_forward) and a backward that leaves the per-edge gradients in the [E, F] workspace
    (their source-side reduction is part of dgn_pair_gather_backward)."""
    kind, n_graphs, F, aggs, scs, k_used = CASES[name]
    if folded:
        scs = "identity"
    base = make_samples(kind, n_graphs, seed=0)
    samples = base * scale                                   # replicated batch: same statistics, `scale` x the size
    g, _ = collate(samples)
    g.to(dev)
    avg = avg_log_degree(base)
    N, E = g.number_of_nodes(), g.number_of_edges()
    agg_list, sc_list = [AGGREGATORS[a] for a in aggs.split()], [SCALERS[s] for s in scs.split()]
    spec = AggSpec(agg_list, sc_list, avg, F, g.ndata["eig"].shape[1])
    A, S = len(agg_list), spec.S
    W = F + S * A * F
    per_set = 4 * N * (2 * W + 7 * F) + 4 * E * F
    rot = int(max(2, min(8, (300 << 20) // per_set + 1)))
    gen = torch.Generator(device=dev).manual_seed(0)
    sets = []
    for _ in range(rot):
        t = {k: torch.randn(N, F, device=dev, generator=gen) for k in ("h", "P", "Q")}
        t["out"] = torch.empty(N, W, device=dev)
        t["gy"] = torch.randn(N, W, device=dev, generator=gen)
        t["dP"], t["dQ"], t["dh"] = (torch.empty(N, F, device=dev) for _ in range(3))
        t["ws"] = torch.empty(max(E, 1), F, device=dev)
        sets.append(t)
    eig = g.ndata["eig"]
    tf = time_graph(lambda t: agg_forward_raw(g, spec, _lib.MSG_AFFINE, t["P"], t["Q"], None, t["h"], eig, t["out"], True), sets)
    tb = time_graph(lambda t: agg_backward_raw(g, spec, _lib.MSG_AFFINE, t["P"], t["Q"], None, t["h"], eig, t["gy"], True,
                                               d_x=None if folded else t["dP"], d_q=t["dQ"], d_h=t["dh"],
                                               edge_ws=t["ws"]), sets)
    # the per-batch eigen-field build (one launch shared by all layers of a step), timed on its own
    t_field = 0.0
    if ops.FIELD_ENABLED:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            g.invalidate_fields()
            g.field(spec, eig)
        b.record()
        torch.cuda.synchronize()
        t_field = a.elapsed_time(b) * 1e-3 / 10
    bf = 4 * (E + N * (3 * F + k_used + 1) + N * S * A * F)
    bb = bf + 4 * N * 3 * F
    if folded:                                               # d_P is not produced; the [E, F] spill is written once
        bb += 4 * E * F - 4 * N * F
    pk = peak()
    return {"case": name, "scale": scale, "layout": "folded" if folded else "reference", "tile": os.environ.get("DGN_TILE", "0"), "graphs": len(samples), "N": N, "E": E, "F": F, "A": A, "S": S, "rot": rot,
            "fwd_us": tf * 1e6, "bwd_us": tb * 1e6, "field_build_us": t_field * 1e6, "bytes_fwd": bf, "bytes_bwd": bb,
            "fwd_gbs": bf / tf / 1e9, "bwd_gbs": bb / tb / 1e9, "fwd_frac": bf / tf / 1e9 / pk,
            "bwd_frac": bb / tb / 1e9 / pk, "frac": (bf + bb) / (tf + tb) / 1e9 / pk, "peak_gbs": pk,
            "edges_per_s_fwd_bwd": E / (tf + tb)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="zinc,pattern,cifar,molhiv")
    ap.add_argument("--scales", default="1,4,16,64")
    ap.add_argument("--folded", action="store_true", help="raw-aggregate layout of the fused layer (scalers in the GEMM)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    for name in args.cases.split(","):
        for sc in [int(s) for s in args.scales.split(",")]:
            if name == "pattern" and sc > 4:
                continue                                     # 1.5 M edges x scale: keep the sweep bounded
            print(json.dumps(run_case(name, sc, dev, args.folded)), flush=True)
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
