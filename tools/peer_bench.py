#!/usr/bin/env python
"""Gradient exchange + optimizer alone, per step: dgn_allreduce_adam (one launch over NVLink peer memory, one-shot and
two-shot) against NCCL all-reduce + dgn_adam_step, on a flat buffer of the bench model's size.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/peer_bench.py [--numel 546304] [--iters 300]
Device time (CUDA events) of `iters` back-to-back exchanges, max over ranks; rank 0 prints one JSON line."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def timed(fn, iters, dev):
    for _ in range(20):
        fn()
    torch.cuda.synchronize(dev)
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([a.elapsed_time(b) * 1e3 / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--numel", type=int, default=546304)
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from dgn_b200.engine import FlatAdam
    from dgn_b200.parallel import PeerGradients, allreduce_sum_
    world = dist.get_world_size()
    out = {"numel": args.numel, "world": world, "unit": "us per exchange+update", "grid": os.environ.get("DGN_AR_GRID", "auto"),
           "tag": args.tag}
    for name, env in (("peer_one_shot", str(world)), ("peer_two_shot", "0")):
        os.environ["DGN_AR_ONESHOT_MAX_WORLD"] = env
        peer = PeerGradients(args.numel, dev)
        p = torch.randn(args.numel, device=dev)
        peer.grad.normal_()
        opt = FlatAdam(p, peer.grad, lr=1e-3, grad_scale=1.0 / world)
        out[name] = timed(lambda: peer.allreduce_adam(opt), args.iters, dev)
        assert not peer.timed_out()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10):
                peer.allreduce_adam(opt)
        out[name + "_graphed"] = timed(g.replay, args.iters // 10, dev) / 10
    p = torch.randn(args.numel, device=dev)
    gbuf = torch.randn(args.numel, device=dev)
    opt = FlatAdam(p, gbuf, lr=1e-3, grad_scale=1.0 / world)

    def nccl():
        allreduce_sum_(gbuf)
        opt.step()
    out["nccl_allreduce_plus_adam"] = timed(nccl, args.iters, dev)
    out["adam_only"] = timed(opt.step, args.iters, dev)
    if dist.get_rank() == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
