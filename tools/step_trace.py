#!/usr/bin/env python
"""In-graph kernel timeline of the bench step (CUPTI through torch.profiler): per-kernel durations as they are INSIDE
the replayed CUDA graph (warm L2, real overlap with the side stream), idle gaps, and the critical path per stream.

    python tools/step_trace.py [--steps 5] [--out gpurun_out/step_trace.json]

ncu launch lists (profiles/*launches*) are cold-cache and serialised; this is the complementary view."""
import argparse
import collections
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--workload", default="zinc")
    ap.add_argument("--hidden", type=int, default=0)
    ap.add_argument("--aggregators", default="")
    ap.add_argument("--out", default=os.path.join(REPO, "gpurun_out", "step_trace.json"))
    args = ap.parse_args()
    import bench
    from dgn_b200.data.synthetic import make_samples, avg_log_degree
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = False
    w = dict(bench.WORKLOADS[args.workload])
    if args.hidden:
        w["hidden"] = args.hidden
    if args.aggregators:
        w["aggregators"] = args.aggregators
    avg_log = avg_log_degree(make_samples(w["kind"], 1000 if w["kind"] != "pattern" else 64, seed=12345))
    pools = [make_samples(w["kind"], w["graphs_per_gpu"], seed=0)]
    net, step, host_batches, targets_host, _, _ = bench.build_step(w, pools, avg_log, dev, eager=False)
    g, tg = host_batches[0], targets_host[0]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(5):
        step.load(g, tg)
        step.run()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(args.steps):
            flush.zero_()
            step.load(g, tg)
            step.run()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name, getattr(e, "stream", None) or 0) for e in evs),
                key=lambda t: t[0])
    # split into steps at the L2-flush fill kernels
    steps, cur = [], []
    for k in ks:
        if "FillFunctor<unsigned char>" in k[2]:
            if cur:
                steps.append(cur)
            cur = []
        elif "Memcpy" not in k[2] and "Memset" not in k[2]:
            cur.append(k)
    if cur:
        steps.append(cur)
    steps = [s for s in steps if len(s) > 20]
    agg = collections.OrderedDict()
    span, busy = [], []
    for st in steps:
        t0, t1 = st[0][0], max(k[1] for k in st)
        span.append(t1 - t0)
        for a, b, name, stream in st:
            c = agg.setdefault(name[:90], [0, 0.0])
            c[0] += 1
            c[1] += b - a
    n = len(steps)
    print("steps traced: %d, span per step (first kernel start -> last kernel end): %.1f us" % (n, sum(span) / max(n, 1)))
    print("| per step | avg us | total us / step | kernel |\n|---:|---:|---:|---|")
    for name, (cnt, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %.1f | %.2f | %.1f | `%s` |" % (cnt / n, tot / cnt, tot / n, name))
    # timeline of the last step: start offsets, durations, gap to the previous kernel end on any stream
    st = steps[-1]
    t0 = st[0][0]
    rows = [{"start_us": a - t0, "dur_us": b - a, "name": name[:60], "stream": stream} for a, b, name, stream in st]
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"span_us": span, "timeline": rows}, open(args.out, "w"))
    prev_end = 0.0
    print("\nlast step timeline (start, dur, gap-after-previous-end, stream, kernel):")
    for r in rows:
        print("%8.1f %6.1f %6.1f  s%-3s %s" % (r["start_us"], r["dur_us"], r["start_us"] - prev_end, str(r["stream"])[-3:], r["name"]))
        prev_end = max(prev_end, r["start_us"] + r["dur_us"])


if __name__ == "__main__":
    main()
