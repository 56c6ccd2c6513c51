#!/usr/bin/env python
"""Benchmark of the DGN hot path: fwd+bwd M-edges/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload zinc|molhiv|pattern|cifar] [--scaling weak|strong]

One "step" = zero_grad + DGNNet forward + loss + backward (+ gradient all-reduce for N>1) + Adam update on one synthetic
mini-batch.  Default workload = BASELINE configs[1]: ZINC-like, 128 graphs per GPU, DGN complex, L=4, hidden 64,
10 aggregators x 3 scalers, k=2 eigenvectors.  The other workloads are BASELINE configs[2..4] (`--workload`):
cifar (configs[2]), molhiv 4 towers (configs[3], global batch 512), pattern (configs[4], global batch 256).
`--scaling weak` (default): the per-GPU batch is fixed (zinc 128, cifar 128, molhiv 64, pattern 32 graphs per GPU);
`--scaling strong`: the GLOBAL batch is fixed (zinc 128, molhiv 512, pattern 256) and sharded over the ranks.
For N>1 the default line also carries a `strong` object (the global-batch-128 figure of the metric's wording).
Prints ONE JSON line (rank 0).

* value      device-resident inputs, CUDA-event time of the K steps (L2 flushed between steps), max over ranks
* e2e        same metric with HOST inputs through the public API (TrainStep.load_ids / run_logged): per step one H2D
             copy of the sampler's index list from pinned memory, the batch is collated on the DEVICE from the
             HBM-resident dataset fragments inside the timed region, and the loss is copied to pinned host memory; the
             host reads the loss of step i after queueing step i+1.  Timed like `value` (L2 flush between steps, one
             event pair per step).  `with_blocking_loss_read_every_step_no_l2_flush` = loss.item() after every step
             (round-1 method: one event pair around the loop), `with_host_collated_batches` = round-1 input path
* roofline   the fused aggregation kernels of one layer of the workload as the step runs them (raw aggregates, the
             scalers are folded into the posttrans GEMM), timed alone with CUDA events (CUDA graph of 8 launches over
             rotating operand sets > L2); algorithmic bytes of SURVEY.md 8(d) / DESIGN.md over the measured HBM copy
             peak (MEASURED_PEAKS.json).  `reference_layout` = the same kernels writing / reading the reference's
             [N, S*A*F] reduce_func layout (dgn_agg_forward / dgn_agg_backward as the ABI exposes them), `at_scale` = the
             same workload replicated 16x (one launch is a single latency-bound wave at 128 graphs)
* cpu_baseline / --impl reference   the reference's own python path on the host cores: the UNMODIFIED
             realworld_benchmark/nets modules byte-compiled into oracle/_ref (kind "reference"), else the oracle port
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

POOL = 8                      # distinct batches cycled through
UNIT = "M-edges/s"
S3 = "identity amplification attenuation"

# BASELINE.json configs[1..4] (SURVEY.md 8(d) for the aggregator strings)
WORKLOADS = {
    "zinc": dict(kind="zinc", metric="DGN fwd+bwd M-edges/sec on ZINC b=128", net="zinc", graphs_per_gpu=128,
                 global_batch=128, hidden=64, L=4, type_net="complex", towers=None,
                 aggregators="mean max min std dir1-dx dir2-dx dir1-dx-no-abs dir2-dx-no-abs dir1-av dir2-av", scalers=S3,
                 text="ZINC-like synthetic, DGN complex L=4 hidden=64, 10 aggregators x 3 scalers, k=2 eigvecs "
                      "(BASELINE configs[1])"),
    "cifar": dict(kind="cifar", metric="DGN fwd+bwd M-edges/sec on CIFAR10-superpixel b=128", net="cifar",
                  graphs_per_gpu=128, global_batch=128, hidden=64, L=4, type_net="complex", towers=None,
                  aggregators="mean dir1-dx dir2-dx", scalers="identity",
                  text="CIFAR10-superpixel-like synthetic (directed 8-NN, ~118 nodes), DGN complex L=4 hidden=64, "
                       "mean+dir1-dx+dir2-dx, k=2 (BASELINE configs[2])"),
    "molhiv": dict(kind="molhiv", metric="DGN fwd+bwd M-edges/sec on ogbg-molhiv b=512", net="hiv", graphs_per_gpu=64,
                   global_batch=512, hidden=80, L=4, type_net="towers", towers=4,
                   aggregators="mean max min dir1-dx dir2-dx dir1-av dir2-av", scalers="identity",
                   text="ogbg-molhiv-like synthetic, DGN 4 towers L=4 hidden=80, 7 aggregators, k=2 "
                        "(BASELINE configs[3])"),
    "pattern": dict(kind="pattern", metric="DGN fwd+bwd M-edges/sec on SBM-PATTERN b=256", net="sbm", graphs_per_gpu=32,
                    global_batch=256, hidden=48, L=4, type_net="complex", towers=None,
                    aggregators="mean dir1-dx dir2-dx dir3-dx dir4-dx", scalers=S3,
                    text="SBM-PATTERN-like synthetic (100-180 nodes, mean degree ~51), DGN complex L=4 hidden=48, "
                         "mean + dir1..4-dx x 3 scalers, k=4 (BASELINE configs[4])"),
}


def load_synthetic():
    """The numpy-only generators of dgn_b200/data/synthetic.py WITHOUT importing the package (whose __init__ loads
    libdgn_b200.so): the reference arm must not map the product's library."""
    spec = importlib.util.spec_from_file_location("dgn_synthetic", os.path.join(REPO, "dgn_b200", "data", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def net_params(w, avg_log, device):
    import torch
    p = dict(hidden_dim=w["hidden"], out_dim=w["hidden"], in_feat_dropout=0.0, dropout=0.0, L=w["L"],
             type_net=w["type_net"], pos_enc_dim=0, readout="mean", graph_norm=True, batch_norm=True,
             aggregators=w["aggregators"], scalers=w["scalers"], avg_d={"log": torch.tensor(float(avg_log))},
             residual=True, edge_feat=False, edge_dim=0, pretrans_layers=1, posttrans_layers=1, device=device)
    if w["net"] == "zinc":
        p.update(num_atom_type=28, num_bond_type=4)
    elif w["net"] == "hiv":
        p.update(towers=w["towers"])
    elif w["net"] == "sbm":
        p.update(in_dim=3, n_classes=2)
    elif w["net"] == "cifar":
        p.update(in_dim=5, n_classes=10, in_dim_edge=1)
    return p


def workload_config(w, name, n_gpus, scaling):
    per = w["graphs_per_gpu"] if scaling == "weak" else w["global_batch"] // n_gpus
    return {"workload": "%s, batch=%d graphs/GPU" % (w["text"], per), "name": name,
            "global_batch": per * n_gpus, "graphs_per_gpu": per, "aggregators": w["aggregators"], "scalers": w["scalers"],
            "step": "zero_grad+fwd+loss+bwd+adam; GPU arm: whole step replayed from one CUDA graph (padded batch layout)",
            "parallelism": "dp%d" % n_gpus if n_gpus > 1 else "single",
            "l2": "256 MiB buffer written between timed steps (L2 flush)", "batch_pool": POOL}


def make_pools(syn, w, per_gpu, rank, world, scaling):
    """POOL lists of samples for this rank.  weak: every rank draws its own per_gpu graphs; strong: a global batch is
    drawn once (same seed on every rank) and cut into contiguous shards of equal graph count."""
    pools = []
    for b in range(POOL):
        if scaling == "strong" and world > 1:
            glob = syn.make_samples(w["kind"], per_gpu * world, seed=7000 + b)
            pools.append(glob[rank * per_gpu:(rank + 1) * per_gpu])
        else:
            pools.append(syn.make_samples(w["kind"], per_gpu, seed=1000 * rank + b))
    return pools


def targets_of(w, samples):
    import torch
    if w["net"] == "zinc":
        return torch.tensor([float(s["label"]) for s in samples]).unsqueeze(1)
    if w["net"] == "hiv":
        return torch.tensor([float(s["label"]) for s in samples])
    if w["net"] == "sbm":
        return torch.from_numpy(np.concatenate([s["label"] for s in samples]).astype(np.int64))
    return torch.tensor([int(s["label"]) for s in samples], dtype=torch.int64)


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own python path (oracle/_ref = byte-compiled UNMODIFIED rb/nets), else the oracle port
# ------------------------------------------------------------------------------------------------
def cpu_reference_time(w, steps, warmup, threads=None, budget_s=25.0, optimizer=True):
    import torch
    from oracle.build_ref import import_ref, ref_available
    from oracle.graphs import collate_standin
    syn = load_synthetic()
    threads = threads or (os.cpu_count() or 1)
    torch.set_num_threads(threads)
    avg = syn.avg_log_degree(syn.make_samples(w["kind"], 1000 if w["kind"] != "pattern" else 64, seed=12345))
    pools = make_pools(syn, w, w["graphs_per_gpu"], 0, 1, "weak")
    kind = "port"
    if ref_available():
        import_ref()
        mod = {"zinc": "molecules_graph_regression", "hiv": "HIV_graph_classification", "sbm": "SBMs_node_classification",
               "cifar": "superpixels_graph_classification"}[w["net"]]
        Net = importlib.import_module("nets.%s.dgn_net" % mod).DGNNet
        kind = "reference"
    else:
        from oracle import task_nets
        Net = {"zinc": task_nets.ZincNet, "hiv": task_nets.HivNet, "sbm": task_nets.PatternNet,
               "cifar": task_nets.SuperpixelNet}[w["net"]]
    torch.manual_seed(41)
    net = Net(net_params(w, avg, "cpu")).train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=3e-6)
    batches = []
    for samples in pools:
        g, _, snorm_n, snorm_e = collate_standin(samples)
        batches.append((g, g.ndata["feat"], g.edata["feat"], snorm_n, snorm_e, targets_of(w, samples),
                        g.number_of_edges()))

    def loss_of(scores, tgt):
        if w["net"] == "hiv" and kind == "reference":     # the reference's loss moves the labels to 'cuda' (dgn_net.py:88)
            return torch.nn.BCEWithLogitsLoss()(scores, tgt.float().unsqueeze(-1))
        return net.loss(scores, tgt)

    def step(i):
        g, x, e, sn, se, tgt, _ = batches[i % POOL]
        opt.zero_grad()
        loss = loss_of(net(g, x, e, sn, se), tgt)
        loss.backward()
        if optimizer:
            opt.step()
        return loss

    for i in range(max(warmup, 1)):
        step(i)
    times, edges, t_all = [], 0, time.perf_counter()
    for i in range(steps):
        t0 = time.perf_counter()
        step(i)
        times.append(time.perf_counter() - t0)
        edges += batches[i % POOL][6]
        if time.perf_counter() - t_all > budget_s:
            break
    total = float(np.sum(times))
    return {"ms_per_step": 1e3 * total / len(times), "value": edges / total / 1e6, "edges": edges / len(times),
            "steps_done": len(times), "cores": torch.get_num_threads(), "kind": kind}


def cpu_sample_text(w, r):
    src = ("the UNMODIFIED realworld_benchmark/nets modules (byte-compiled into oracle/_ref)" if r["kind"] == "reference"
           else "the oracle port of realworld_benchmark/nets")
    return ("%d full steps over a pool of %d %s batches of %d graphs (%.0f edges/step) of %s on the DGL-0.4.2 stand-in, "
            "%.1f ms/step" % (r["steps_done"], POOL, w["kind"], w["graphs_per_gpu"], r["edges"], src, r["ms_per_step"]))


def cpu_leg(workload, steps, warmup, budget, threads=0, no_opt=False):
    """cpu_reference_time in a child process with CUDA hidden; returns its dict or None."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload, "--steps", str(steps),
           "--warmup", str(warmup), "--cpu-leg", "--cpu-budget", str(budget), "--cpu-threads", str(threads)]
    if no_opt:
        cmd.append("--cpu-no-opt")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    try:
        out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=budget * 6 + 120).stdout
        return json.loads(out.strip().splitlines()[-1])
    except Exception:
        return None


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    if args.cpu_leg:
        r = cpu_reference_time(w, args.steps, args.warmup, threads=args.cpu_threads or None, budget_s=args.cpu_budget,
                               optimizer=not args.cpu_no_opt)
        print(json.dumps(r), flush=True)
        return
    r = cpu_reference_time(w, args.steps, args.warmup, budget_s=150.0)
    # hygiene: this arm must not have touched the product - neither the package nor its library
    assert not any(m == "dgn_b200" or m.startswith("dgn_b200.") for m in sys.modules), "reference arm imported dgn_b200"
    try:
        assert "libdgn_b200" not in open("/proc/self/maps").read(), "reference arm mapped libdgn_b200.so"
    except OSError:
        pass
    line = {"impl": "reference", "metric": w["metric"], "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": r["steps_done"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # the SAME config dict as the GPU arm prints for these flags; a CPU step is a bounded sample of it: one
            # rank's shard (the metric, edges / s, does not depend on how many shards are timed)
            "config": workload_config(w, args.workload, max(args.gpus, 1), args.scaling),
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                             "sample": cpu_sample_text(w, r) + ("" if args.gpus <= 1 else
                                                                "; one rank's shard of the %d-GPU workload per step" % args.gpus)},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, smax = [], set(), None
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    smax = float(f[2])
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm))
        return out


def agg_bytes(N, E, F, A, S, r_ops, k_used):
    """Algorithmic HBM bytes of one fused aggregation (GATHER mode), fwd and bwd."""
    fwd = 4 * (E + N * (r_ops * F + k_used + 1) + N * S * A * F)
    bwd = fwd + 4 * N * r_ops * F
    return fwd, bwd


def kernel_roofline(w, graph, avg_log, device, folded, rot=8, replays=10):
    """Times dgn_agg_forward / dgn_agg_backward alone on the layer operands of the workload.

    ``folded``: raw aggregates (one scaler) as the step runs them, else the reference's [N, S*A*F] layout.  ``rot``
    operand sets are rotated inside one captured CUDA graph so that every launch finds its inputs cold and no CPU launch
    gap sits between the CUDA events; the reported time is (event time of the replays) / (replays * rot).  The backward
    leaves the per-edge gradients in the workspace (their source-side reduction is part of the pretrans kernel)."""
    import torch
    from dgn_b200 import _lib
    from dgn_b200.nets.aggregators import AGGREGATORS
    from dgn_b200.nets.scalers import SCALERS as SC
    from dgn_b200.ops import AggSpec, agg_forward_raw, agg_backward_raw
    tw = w["towers"] or 1
    N, E, F = graph.number_of_nodes(), graph.number_of_edges(), w["hidden"] // tw
    n_real, e_real = graph.n_real_nodes, graph.n_real_edges
    aggs = [AGGREGATORS[a] for a in w["aggregators"].split()]
    scal = ["identity"] if folded else w["scalers"].split()
    spec = AggSpec(aggs, [SC[s] for s in scal], avg_log, F, graph.ndata["eig"].shape[1])
    A, S = len(aggs), spec.S
    simple = w["type_net"] == "simple"
    Wd = (0 if simple else F) + S * A * F
    gen = torch.Generator(device=device).manual_seed(0)
    eig = graph.ndata["eig"]
    sets = []
    for _ in range(rot):
        t = {k: torch.randn(N, F, device=device, generator=gen) for k in ("h", "P", "Q")}
        t["out"] = torch.empty(N, Wd, device=device)
        t["gy"] = torch.randn(N, Wd, device=device, generator=gen)
        t["dP"], t["dQ"], t["dh"] = (torch.empty(N, F, device=device) for _ in range(3))
        t["ws"] = torch.empty(max(E, 1), F, device=device)
        sets.append(t)

    def fwd(t):
        agg_forward_raw(graph, spec, _lib.MSG_AFFINE, t["P"], t["Q"], None, t["h"], eig, t["out"], True)

    def bwd(t):
        agg_backward_raw(graph, spec, _lib.MSG_AFFINE, t["P"], t["Q"], None, t["h"], eig, t["gy"], True,
                         d_x=None if folded else t["dP"], d_q=t["dQ"], d_h=t["dh"], edge_ws=t["ws"])

    def timed(fn):
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            fn(sets[0])
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize()
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg):
            for t in sets:
                fn(t)
        for _ in range(3):
            cg.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(replays):
            cg.replay()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e-3 / (replays * rot)

    t_f, t_b = timed(fwd), timed(bwd)
    k_used = len({a.eig_idx for a in aggs if a.kind >= _lib.AGG_DIR_AV})
    bf, bb = agg_bytes(n_real, e_real, F, A, S, 3, k_used)
    if folded:                                    # d_P is not produced; + the [E, F] spill the pretrans kernel reduces
        bb += 4 * e_real * F - 4 * n_real * F
    return {"fwd_us": t_f * 1e6, "bwd_us": t_b * 1e6, "bytes_fwd": bf, "bytes_bwd": bb,
            "achieved": (bf + bb) / (t_f + t_b) / 1e9, "fwd_gbs": bf / t_f / 1e9, "bwd_gbs": bb / t_b / 1e9,
            "medges_per_s": e_real / (t_f + t_b) / 1e6}          # edges / kernel time of one layer


def measured_traffic():
    """DRAM bytes of one forward + backward launch from the committed ncu capture (profiles/), or None."""
    for name in ("r2_agg_traffic.json", "r1_agg_traffic.json"):
        try:
            t = json.load(open(os.path.join(REPO, "profiles", name)))
            return int(t["fwd_bytes"]) + int(t["bwd_bytes"]), name
        except Exception:
            continue
    return None, None


def measured_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def build_step(w, pools, avg_log, dev, eager, lr=1e-3):
    """Network + TrainStep + the pool as packed pinned host batches (and their targets)."""
    import torch
    from dgn_b200.engine import TrainStep
    from dgn_b200.graph import collate
    from dgn_b200.task_nets import (HIV_graph_classification, SBMs_node_classification, molecules_graph_regression,
                                    superpixels_graph_classification)
    Net = {"zinc": molecules_graph_regression.DGNNet, "hiv": HIV_graph_classification.DGNNet,
           "sbm": SBMs_node_classification.DGNNet, "cifar": superpixels_graph_classification.DGNNet}[w["net"]]
    cap_n = (int(max(sum(s["n"] for s in p) for p in pools) * 1.03) + 63) // 64 * 64
    cap_e = (int(max(sum(len(s["src"]) for s in p) for p in pools) * 1.03) + 63) // 64 * 64
    capacity = None if eager else (cap_n, cap_e)
    host_batches, targets_host = [], []
    for samples in pools:
        g, _ = collate(samples, capacity=capacity)
        host_batches.append(g)
        t = targets_of(w, samples)
        targets_host.append(t.pin_memory() if torch.cuda.is_available() else t)
    torch.manual_seed(41)
    net = Net(net_params(w, avg_log, dev)).to(dev).train()
    template, _ = collate(pools[0], capacity=capacity)         # its device views become the static batch buffers
    step = TrainStep(net, template, targets_host[0], lr=lr, weight_decay=3e-6, graphed=not eager)
    dataset = None
    if not eager:
        # the pool as a dataset resident in HBM (per-graph fragments + targets): a step's host input is an index list
        from dgn_b200.data.device_dataset import DeviceDataset
        flat = [s for p in pools for s in p]
        dataset = DeviceDataset(flat, dev, targets=targets_of(w, flat))
    return net, step, host_batches, targets_host, capacity, dataset


def time_steps(step, host_batches, targets_host, edges, dev, args, barrier, eager, dataset=None):
    """(device-resident ms, e2e ms, edges, launches, h2d, d2h, e2e-with-host-collated-batches ms) of `args.steps` steps."""
    import torch
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    if eager:        # eager batches differ in size: re-bind the step's graph object per batch
        from dgn_b200.graph import collate  # noqa: F401
        dev_batches = [None] * POOL
        dev_targets = [t.to(dev) for t in targets_host]

        def stage_host(i):
            step.g = host_batches[i % POOL].to(dev)
            step.targets = targets_host[i % POOL].to(dev, non_blocking=True)

        def stage_dev(i):
            if dev_batches[i % POOL] is None:
                dev_batches[i % POOL] = host_batches[i % POOL].to(dev)
            step.g, step.targets = dev_batches[i % POOL], dev_targets[i % POOL]
    else:
        dev_blobs = [g._host_blob.to(dev) for g in host_batches]
        dev_targets = [t.to(dev) for t in targets_host]

        def stage_host(i):
            step.load(host_batches[i % POOL], targets_host[i % POOL])

        def stage_dev(i):
            step.load_device(dev_blobs[i % POOL], dev_targets[i % POOL])
    torch.cuda.synchronize()
    for i in range(args.warmup):
        stage_dev(i)
        step.run()
    barrier()
    evs, n_edges = [], 0
    for i in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        stage_dev(i)
        step.run()
        b.record()
        evs.append((a, b))
        n_edges += edges[i % POOL]
    barrier()
    step_ms = sum(a.elapsed_time(b) for a, b in evs)
    launches = step.launches_per_step * args.steps
    # end-to-end (a): host-collated packed batches in, loss out (round-1 definition; collation NOT in the timed region)
    for i in range(min(args.warmup, 3)):
        stage_host(i)
        step.run().item()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h2d = d2h = 0
    e0.record()
    for i in range(args.steps):
        stage_host(i)                                            # ONE packed H2D copy from pinned memory (+ targets)
        step.run().item()                                        # D2H read of the step's loss
        h2d += host_batches[i % POOL].h2d_bytes + targets_host[i % POOL].numel() * targets_host[i % POOL].element_size()
        d2h += 4
    e1.record()
    barrier()
    host_ms = e0.elapsed_time(e1)
    if dataset is None:
        return step_ms, host_ms, n_edges, launches, h2d, d2h, host_ms, None
    # end-to-end (b), the headline: the sampler's index list in (pinned host memory -> H2D), the batch is COLLATED ON THE
    # DEVICE from the dataset-resident fragments inside the timed region, loss out
    B = host_batches[0].graph_capacity
    ids = [torch.arange(b * B, (b + 1) * B, dtype=torch.int32).pin_memory() for b in range(POOL)]
    for i in range(min(args.warmup, 3)):
        step.load_ids(dataset, ids[i % POOL])
        step.run().item()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h2d = d2h = 0
    e0.record()
    for i in range(args.steps):
        step.load_ids(dataset, ids[i % POOL])                    # 4 B per graph H2D + one collation launch
        step.run().item()
        h2d += ids[i % POOL].numel() * 4
        d2h += 4
    e1.record()
    barrier()
    sync_ms = e0.elapsed_time(e1)
    # the same loop with the loss of step i read by the host after step i + 1 has been queued (TrainStep.run_logged: async
    # 4-byte D2H into a pinned ring): every step still copies its inputs in and its loss out inside the timed region
    # Timed like `value`: the L2 is flushed between steps and every step has its own event pair around
    # [index H2D -> collation -> step -> loss D2H]; a host that falls behind shows up as GPU idle time inside the pair.
    pending, seen, evs = None, 0.0, []
    for i in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step.load_ids(dataset, ids[i % POOL])
        cur = step.run_logged()
        b.record()
        evs.append((a, b))
        if pending is not None:
            pending[1].synchronize()
            seen += float(pending[0])
        pending = cur
    pending[1].synchronize()
    seen += float(pending[0])
    barrier()
    piped_ms = sum(a.elapsed_time(b) for a, b in evs)
    if not math.isfinite(seen):
        raise SystemExit("bench.py: non-finite loss in the e2e leg")
    return step_ms, sync_ms, n_edges, launches, h2d, d2h, host_ms, piped_ms


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from dgn_b200.data import synthetic as syn

    w = WORKLOADS[args.workload]
    eager = args.eager or w["net"] == "sbm"       # node-level targets (class-weighted loss over the real nodes): unpadded
    avg_log = syn.avg_log_degree(syn.make_samples(w["kind"], 1000 if w["kind"] != "pattern" else 64, seed=12345))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(scaling):
        per = w["graphs_per_gpu"] if scaling == "weak" else max(w["global_batch"] // world, 1)
        pools = make_pools(syn, w, per, rank, world, scaling)
        net, step, host_batches, targets_host, capacity, dataset = build_step(w, pools, avg_log, dev, eager)
        edges = [g.n_real_edges for g in host_batches]
        step_ms, sync_ms, n_edges, launches, h2d, d2h, host_ms, piped_ms = time_steps(
            step, host_batches, targets_host, edges, dev, args, barrier, eager, dataset)
        if getattr(step, "peer", None) is not None and step.peer.timed_out():
            raise SystemExit("bench.py: a rank did not reach a barrier of dgn_allreduce_adam - results invalid")
        stats = torch.tensor([step_ms, sync_ms, float(n_edges), host_ms, piped_ms or sync_ms], device=dev,
                             dtype=torch.float64)
        if world > 1:
            mx, sm = stats.clone(), stats.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            step_ms, sync_ms, total_edges, host_ms, e2e_ms = float(mx[0]), float(mx[1]), float(sm[2]), float(mx[3]), float(mx[4])
        else:
            total_edges, e2e_ms = float(n_edges), float(stats[4])
        return {"value": total_edges / (step_ms * 1e-3) / 1e6, "e2e": total_edges / (e2e_ms * 1e-3) / 1e6,
                "e2e_sync": total_edges / (sync_ms * 1e-3) / 1e6, "pipelined": piped_ms is not None,
                "ms_per_step": step_ms / args.steps, "e2e_ms_per_step": e2e_ms / args.steps, "launches": launches,
                "e2e_host": total_edges / (host_ms * 1e-3) / 1e6, "device_collate": dataset is not None,
                "h2d": h2d // args.steps, "d2h": d2h // args.steps, "edges_per_step_per_gpu": float(np.mean(edges)),
                "params": int(sum(p.numel() for p in net.parameters())), "step": step, "pools": pools,
                "graphs_per_gpu": per}

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_wall = time.perf_counter()
    main = measure(args.scaling)
    strong = None
    if world > 1 and args.scaling == "weak" and w["global_batch"] // world >= 1 and not args.no_strong:
        step_keep = main.pop("step")              # keep the weak step alive for the roofline leg
        strong = measure("strong")
        strong.pop("step")
        strong.pop("pools")
        main["step"] = step_keep
    wall_s = time.perf_counter() - t_wall
    clocks = sampler.stop() if rank == 0 else None

    roof = cpu = None
    if rank == 0 and w["hidden"] % 4 == 0:
        from dgn_b200.graph import collate
        step = main["step"]
        peak, peak_src = measured_peak()
        kr = kernel_roofline(w, step.g, avg_log, dev, folded=True)
        kref = kernel_roofline(w, step.g, avg_log, dev, folded=False)
        big, _ = collate(main["pools"][0] * (16 if w["kind"] != "pattern" else 2))
        big.to(dev)
        ks = kernel_roofline(w, big, avg_log, dev, folded=True, rot=2, replays=5)
        del big
        torch.cuda.empty_cache()
        traffic, tsrc = measured_traffic()

        def pack(k):
            return {"achieved": k["achieved"], "frac": k["achieved"] / peak, "fwd_us": k["fwd_us"], "bwd_us": k["bwd_us"],
                    "fwd_frac": k["fwd_gbs"] / peak, "bwd_frac": k["bwd_gbs"] / peak, "bytes_fwd": k["bytes_fwd"],
                    "bytes_bwd": k["bytes_bwd"], "kernel_medges_per_s_per_layer": k["medges_per_s"]}
        roof = {"bound": "hbm", "peak": peak, "unit": "GB/s", "traffic": traffic, "peak_source": peak_src,
                "traffic_note": "dram__bytes_read+write of one fwd+bwd launch, ncu --set full (profiles/%s)" % tsrc,
                "kernel": "dgn::agg_fwd_row_kernel + dgn::agg_bwd_row_kernel = dgn_agg_forward + dgn_agg_backward of one DGN "
                          "layer of the workload exactly as the step launches them (raw aggregates [N, A*F]: the scalers "
                          "are folded into the tcgen05 posttrans GEMM, so a launch moves ~2.4x fewer bytes than the "
                          "reference layout and is a single latency-bound wave at 128 graphs), timed alone: CUDA graph of "
                          "8 launches on rotating operand sets > L2; the per-batch dgn_field_build launch (shared by the "
                          "2*L aggregation launches of a step) is not included",
                "reference_layout": dict(pack(kref), note="same kernels on the reference's reduce_func layout [N, S*A*F] "
                                                          "(what dgn_agg_forward / dgn_agg_backward expose for a caller "
                                                          "that wants rb/nets/dgn_layer.py:94-96 literally; incl. the "
                                                          "source-side reduction launch)"),
                "at_scale": dict(pack(ks), workload="same batch replicated %dx in one launch (folded layout)"
                                 % (16 if w["kind"] != "pattern" else 2))}
        roof.update(pack(kr))
        if world == 1 and not args.no_cpu:
            # separate processes with the GPUs hidden: the unmodified reference moves tensors to 'cuda' when it sees one
            c = cpu_leg(args.workload, steps=60, warmup=2, budget=15.0)
            if c is not None:
                cpu = {"value": c["value"], "unit": UNIT, "cores": c["cores"], "kind": c["kind"],
                       "sample": cpu_sample_text(w, c)}
                c1 = cpu_leg(args.workload, steps=20, warmup=1, budget=6.0, threads=1)
                cn = cpu_leg(args.workload, steps=60, warmup=1, budget=6.0, no_opt=True)
                if c1 is not None:
                    cpu["one_thread"] = {"value": c1["value"], "ms_per_step": c1["ms_per_step"], "cores": 1}
                if cn is not None:
                    cpu["without_optimizer_step"] = {"value": cn["value"], "ms_per_step": cn["ms_per_step"],
                                                     "cores": cn["cores"]}
    if rank == 0:
        cfg = workload_config(w, args.workload, world, args.scaling)
        line = {"metric": w["metric"], "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
                "e2e": {"value": main["e2e"], "unit": UNIT, "h2d_bytes_per_step": main["h2d"],
                        "d2h_bytes_per_step": main["d2h"], "ms_per_step": main["e2e_ms_per_step"],
                        "input": ("sampler index list (pinned host) -> H2D -> batch collated on the device from the "
                                  "HBM-resident dataset, inside the timed region" if main["device_collate"] else
                                  "host-collated packed batch (pinned) -> one H2D copy"),
                        "output": ("every step's loss -> 4-byte D2H into pinned memory; the host reads the loss of step i "
                                   "after it has queued step i+1 (TrainStep.run_logged); L2 flushed between steps, one "
                                   "event pair per step around H2D + collation + step + D2H" if main["pipelined"] else
                                   "loss.item() after every step"),
                        "with_blocking_loss_read_every_step_no_l2_flush": main["e2e_sync"],
                        "with_host_collated_batches": main["e2e_host"]},
                "gpu_launches": main["launches"], "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
                "params": main["params"], "edges_per_step_per_gpu": main["edges_per_step_per_gpu"],
                "wall_s_both_legs": wall_s}
        if strong is not None:
            line["strong"] = {"scaling": "strong", "global_batch": strong["graphs_per_gpu"] * world,
                              "graphs_per_gpu": strong["graphs_per_gpu"], "value": strong["value"], "unit": UNIT,
                              "ms_per_step": strong["ms_per_step"], "e2e": strong["e2e"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="dgn_b200", choices=["dgn_b200", "reference"])
    ap.add_argument("--workload", default="zinc", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-strong", action="store_true", help="N>1: skip the extra strong-scaling measurement")
    ap.add_argument("--eager", action="store_true", help="no CUDA-graph capture, unpadded batches (debug)")
    ap.add_argument("--hidden", type=int, default=0, help="override the hidden width (e.g. 45: the shipped ZINC config)")
    ap.add_argument("--aggregators", default="", help="override the aggregator list")
    ap.add_argument("--cpu-leg", action="store_true", help=argparse.SUPPRESS)       # internal: child of the GPU arm
    ap.add_argument("--cpu-budget", type=float, default=20.0, help=argparse.SUPPRESS)
    ap.add_argument("--cpu-threads", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--cpu-no-opt", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.hidden or args.aggregators:
        w = dict(WORKLOADS[args.workload])
        if args.hidden:
            w["hidden"] = args.hidden
        if args.aggregators:
            w["aggregators"] = args.aggregators
        w["text"] += " [overrides: hidden=%d, aggregators=%r]" % (w["hidden"], w["aggregators"])
        WORKLOADS[args.workload] = w
    if args.impl == "reference":
        # the unmodified reference moves tensors to 'cuda' whenever a GPU is visible (rb/nets/dgn_layer.py:82-84):
        # the CPU arm hides the GPUs before torch initialises CUDA
        os.environ["CUDA_VISIBLE_DEVICES"] = ""
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
