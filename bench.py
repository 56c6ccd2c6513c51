#!/usr/bin/env python
"""Benchmark of the DGN hot path: fwd+bwd M-edges/s on ZINC-like batches (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = zero_grad + DGNNet forward + L1 loss + backward (+ gradient all-reduce for N>1) + Adam
update on one synthetic ZINC-like mini-batch of 128 graphs per GPU (BASELINE configs[1]: DGN complex,
L=4, hidden 64, 10 aggregators x 3 scalers, k=2 eigenvectors).  Prints ONE JSON line (rank 0).

* value      device-resident inputs, CUDA-event time of the K steps (L2 flushed between steps), max over ranks
* e2e        same metric with HOST inputs: per step one H2D copy of the packed batch from pinned memory
             and a D2H read of the loss, inside the timed region
* roofline   fused aggregation kernels (forward + backward of one layer of the workload) timed alone with CUDA
             events (CUDA graph of 8 launches over rotating operand sets > L2), algorithmic bytes of SURVEY.md 8(d) /
             DESIGN.md over the measured HBM copy peak (MEASURED_PEAKS.json).  `at_scale` repeats the measurement on
             the same workload replicated 16x (2048 graphs): at 128 graphs one launch is a single wave of ~6 us, i.e.
             launch-latency bound, the 16x figure shows what the kernels reach once the launch is HBM bound
* cpu_baseline / --impl reference   the oracle port of the reference's python path on the host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

AGGS = "mean max min std dir1-dx dir2-dx dir1-dx-no-abs dir2-dx-no-abs dir1-av dir2-av"
SCALERS = "identity amplification attenuation"
HIDDEN, LAYERS, BATCH = 64, 4, 128
POOL = 8                      # distinct pre-collated batches cycled through
METRIC = "DGN fwd+bwd M-edges/sec on ZINC b=128"
UNIT = "M-edges/s"


def net_params(avg_log, device):
    return dict(num_atom_type=28, num_bond_type=4, hidden_dim=HIDDEN, out_dim=HIDDEN, in_feat_dropout=0.0,
                dropout=0.0, L=LAYERS, type_net="complex", pos_enc_dim=0, readout="mean", graph_norm=True,
                batch_norm=True, aggregators=AGGS, scalers=SCALERS, avg_d={"log": torch.tensor(float(avg_log))},
                residual=True, edge_feat=False, edge_dim=0, pretrans_layers=1, posttrans_layers=1, device=device)


def workload_config(n_gpus):
    return {"workload": "ZINC-like synthetic, batch=128 graphs/GPU, DGN complex L=4 hidden=64, "
                        "10 aggregators x 3 scalers, k=2 eigvecs (BASELINE configs[1])",
            "global_batch": BATCH * n_gpus, "graphs_per_gpu": BATCH, "aggregators": AGGS, "scalers": SCALERS,
            "step": "zero_grad+fwd+L1loss+bwd+adam, whole step replayed from one CUDA graph (padded batch layout)", "parallelism": "dp%d" % n_gpus if n_gpus > 1 else "single",
            "l2": "256 MiB buffer written between timed steps (L2 flush)", "batch_pool": POOL}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's python path (degree-bucketed update_all)
# ------------------------------------------------------------------------------------------------
def cpu_reference_time(steps, warmup, seed=0, budget_s=25.0):
    from dgn_b200.data.synthetic import make_samples, avg_log_degree
    from oracle.graphs import collate_standin
    from oracle.task_nets import ZincNet
    torch.set_num_threads(os.cpu_count() or 1)
    samples = make_samples("zinc", BATCH, seed=seed)
    avg = avg_log_degree(samples)
    g, labels, snorm_n, snorm_e = collate_standin(samples)
    torch.manual_seed(41)
    net = ZincNet(net_params(avg, "cpu")).train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=3e-6)
    x, e, tgt = g.ndata["feat"], g.edata["feat"], labels.float().unsqueeze(1)
    E = g.number_of_edges()

    def step():
        opt.zero_grad()
        loss = net.loss(net(g, x, e, snorm_n, snorm_e), tgt)
        loss.backward()
        opt.step()
        return loss

    for _ in range(max(warmup, 1)):
        step()
    times, t_all = [], time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_all > budget_s:
            break
    ms = 1e3 * float(np.mean(times))
    return {"ms_per_step": ms, "value": E / (ms * 1e-3) / 1e6, "edges": E, "steps_done": len(times),
            "cores": torch.get_num_threads()}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_time(args.steps, args.warmup, budget_s=150.0)
    cfg = workload_config(1)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": r["steps_done"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                             "sample": "one ZINC-like batch of 128 graphs (%d edges), %d full steps of the oracle "
                                       "port of realworld_benchmark/nets on the DGL stand-in" % (r["edges"], r["steps_done"])},
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
                                          "--format=csv,noheader,nounits", "-lms", "20"],
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


def kernel_roofline(graph, avg_log, device, rot=8, replays=10, n_real=None, e_real=None):
    """Times dgn_agg_forward / dgn_agg_backward alone on the layer operands of the bench workload.

    ``rot`` operand sets are rotated inside one captured CUDA graph so that every launch finds its
    inputs cold (rot x ~50 MB > the 126 MB L2) and no CPU launch gap sits between the CUDA events;
    the reported time is (event time of the replays) / (replays * rot)."""
    from dgn_b200 import _lib
    from dgn_b200.nets.aggregators import AGGREGATORS
    from dgn_b200.nets.scalers import SCALERS as SC
    from dgn_b200.ops import AggSpec, agg_forward_raw, agg_backward_raw
    N, E, F = graph.number_of_nodes(), graph.number_of_edges(), HIDDEN
    n_real = graph.n_real_nodes if n_real is None else n_real
    e_real = graph.n_real_edges if e_real is None else e_real
    aggs = [AGGREGATORS[a] for a in AGGS.split()]
    spec = AggSpec(aggs, [SC[s] for s in SCALERS.split()], avg_log, F, graph.ndata["eig"].shape[1])
    A, S = len(aggs), 3
    W = F + S * A * F
    gen = torch.Generator(device=device).manual_seed(0)
    eig = graph.ndata["eig"]
    sets = []
    for _ in range(rot):
        t = {k: torch.randn(N, F, device=device, generator=gen) for k in ("h", "P", "Q")}
        t["out"] = torch.empty(N, W, device=device)
        t["gy"] = torch.randn(N, W, device=device, generator=gen)
        t["dP"], t["dQ"], t["dh"] = (torch.empty(N, F, device=device) for _ in range(3))
        t["ws"] = torch.empty(max(E, 1), F, device=device)
        sets.append(t)

    def fwd(t):
        agg_forward_raw(graph, spec, _lib.MSG_AFFINE, t["P"], t["Q"], None, t["h"], eig, t["out"], True)

    def bwd(t):
        agg_backward_raw(graph, spec, _lib.MSG_AFFINE, t["P"], t["Q"], None, t["h"], eig, t["gy"], True,
                         d_x=t["dP"], d_q=t["dQ"], d_h=t["dh"], edge_ws=t["ws"])

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
    bf, bb = agg_bytes(n_real, e_real, F, A, S, 3, 2)
    return {"fwd_us": t_f * 1e6, "bwd_us": t_b * 1e6, "bytes_fwd": bf, "bytes_bwd": bb,
            "achieved_gbs": (bf + bb) / (t_f + t_b) / 1e9, "fwd_gbs": bf / t_f / 1e9, "bwd_gbs": bb / t_b / 1e9,
            "medges_per_s": e_real / (t_f + t_b) / 1e6}          # edges / kernel time of one layer


def measured_traffic():
    """DRAM bytes of one forward + backward launch from the committed ncu capture (profiles/), or None."""
    try:
        t = json.load(open(os.path.join(REPO, "profiles", "r1_agg_traffic.json")))
        return int(t["fwd_bytes"]) + int(t["bwd_bytes"])          # TypeError -> None while a leg is unmeasured
    except Exception:
        return None


def measured_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def run_gpu_arm(args):
    import torch.distributed as dist
    from dgn_b200.data.synthetic import make_samples, avg_log_degree
    from dgn_b200.engine import TrainStep
    from dgn_b200.graph import collate
    from dgn_b200.task_nets.molecules_graph_regression import DGNNet

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

    # ---- workload: POOL batches of 128 ZINC-like graphs per rank, pinned + packed on the host ------------
    ref_samples = make_samples("zinc", 1000, seed=12345)
    avg_log = avg_log_degree(ref_samples)                      # avg_d['log'] over a 1000-graph "training set"
    pools = [make_samples("zinc", BATCH, seed=1000 * rank + b) for b in range(POOL)]
    # one fixed layout for all batches: capacity = largest batch of the pool + ~3 %, multiple of 64
    cap_n = (int(max(sum(s["n"] for s in p) for p in pools) * 1.03) + 63) // 64 * 64
    cap_e = (int(max(sum(len(s["src"]) for s in p) for p in pools) * 1.03) + 63) // 64 * 64
    capacity = (cap_n, cap_e) if not args.eager else None
    host_batches, targets_host = [], []
    for samples in pools:
        g, labels = collate(samples, capacity=capacity)
        host_batches.append(g)
        targets_host.append(labels.float().unsqueeze(1).pin_memory())
    edges = [g.n_real_edges for g in host_batches]

    torch.manual_seed(41)
    net = DGNNet(net_params(avg_log, dev)).to(dev).train()
    n_params = int(sum(p.numel() for p in net.parameters()))
    template, _ = collate(pools[0], capacity=capacity)         # its device views become the static batch buffers
    step = TrainStep(net, template, targets_host[0], lr=1e-3, weight_decay=3e-6, graphed=not args.eager)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.eager:        # eager batches differ in size: re-bind the step's graph object per batch
        def stage_host(i):
            step.g = host_batches[i % POOL].to(dev)
            step.targets = targets_host[i % POOL].to(dev, non_blocking=True)
        dev_graphs = None
    else:
        def stage_host(i):
            step.load(host_batches[i % POOL], targets_host[i % POOL])

    # ---- (1) device-resident arm ---------------------------------------------------------------------------
    if args.eager:
        dev_batches = [collate(p)[0].to(dev) for p in pools]
        dev_targets = [t.to(dev) for t in targets_host]

        def stage_dev(i):
            step.g, step.targets = dev_batches[i % POOL], dev_targets[i % POOL]
    else:
        dev_blobs = [g._host_blob.to(dev) for g in host_batches]
        dev_targets = [t.to(dev) for t in targets_host]

        def stage_dev(i):
            step.load_device(dev_blobs[i % POOL], dev_targets[i % POOL])
    torch.cuda.synchronize()
    for i in range(args.warmup):
        stage_dev(i)
        step.run()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    evs, n_edges = [], 0
    t_wall = time.perf_counter()
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
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    launches = step.launches_per_step * args.steps
    step_ms = sum(a.elapsed_time(b) for a, b in evs)

    # ---- (2) end-to-end arm: host buffers in, loss out ---------------------------------------------------
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
        h2d += host_batches[i % POOL].h2d_bytes + targets_host[i % POOL].numel() * 4
        d2h += 4
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    e2e_ms = e0.elapsed_time(e1)

    # ---- max over ranks, whole-job aggregate ------------------------------------------------------------
    stats = torch.tensor([step_ms, e2e_ms, float(n_edges)], device=dev, dtype=torch.float64)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        step_ms, e2e_ms, total_edges = float(mx[0]), float(mx[1]), float(sm[2])
    else:
        total_edges = float(n_edges)
    value = total_edges / (step_ms * 1e-3) / 1e6
    e2e_value = total_edges / (e2e_ms * 1e-3) / 1e6

    roof = cpu = None
    if rank == 0:
        kr = kernel_roofline(step.g, avg_log, dev)
        peak, peak_src = measured_peak()
        big, _ = collate(pools[0] * 16)                           # the same batch 16 x: 2048 graphs, ~98 k edges
        big.to(dev)
        ks = kernel_roofline(big, avg_log, dev, rot=2, replays=5)
        del big
        torch.cuda.empty_cache()
        roof = {"bound": "hbm", "achieved": kr["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kr["achieved_gbs"] / peak, "traffic": measured_traffic(), "peak_source": peak_src,
                "traffic_note": "dram__bytes_read+write of one fwd+bwd launch at this workload, ncu --set full "
                                "(profiles/r1_agg_traffic.json); the 24 MB of output / gradient stay in the 126 MB L2 for the "
                                "duration of one cold launch, so traffic < algorithmic bytes here; at 16x the batch the "
                                "captures show 361 / 461 MB for 386 / 421 MB algorithmic (no wasted re-reads)",
                "kernel": "dgn::agg_fwd_row_kernel + agg_bwd_row_kernel + agg_bwd_src_kernel = dgn_agg_forward + "
                          "dgn_agg_backward of one DGN layer of the bench workload, timed alone: CUDA graph of 8 launches "
                          "on rotating operand sets > L2; the per-batch dgn_field_build launch (shared by the 8 "
                          "aggregation launches of a step) is not included",
                "fwd_us": kr["fwd_us"], "bwd_us": kr["bwd_us"], "bytes_fwd": kr["bytes_fwd"],
                "bytes_bwd": kr["bytes_bwd"], "fwd_gbs": kr["fwd_gbs"], "bwd_gbs": kr["bwd_gbs"],
                "fwd_frac": kr["fwd_gbs"] / peak, "bwd_frac": kr["bwd_gbs"] / peak,
                "kernel_medges_per_s_per_layer": kr["medges_per_s"],
                "at_scale": {"workload": "same batch replicated 16x (2048 graphs) in one launch",
                             "achieved": ks["achieved_gbs"], "frac": ks["achieved_gbs"] / peak,
                             "fwd_us": ks["fwd_us"], "bwd_us": ks["bwd_us"], "fwd_frac": ks["fwd_gbs"] / peak,
                             "bwd_frac": ks["bwd_gbs"] / peak, "bytes_fwd": ks["bytes_fwd"],
                             "bytes_bwd": ks["bytes_bwd"], "kernel_medges_per_s_per_layer": ks["medges_per_s"]}}
        if world == 1 and not args.no_cpu:
            c = cpu_reference_time(steps=60, warmup=2, budget_s=20.0)
            cpu = {"value": c["value"], "unit": UNIT, "cores": c["cores"], "kind": "port",
                   "sample": "one ZINC-like batch of 128 graphs (%d edges), %d steps of the oracle port on the "
                             "DGL stand-in, %.1f ms/step" % (c["edges"], c["steps_done"], c["ms_per_step"])}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": step_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(workload_config(world), params=n_params,
                               edges_per_step_per_gpu=float(np.mean(edges))),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d // args.steps,
                        "d2h_bytes_per_step": d2h // args.steps, "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
                "wall_ms_per_step_incl_flush": wall_ms / args.steps}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="dgn_b200", choices=["dgn_b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--eager", action="store_true", help="no CUDA-graph capture, unpadded batches (debug)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
