"""torchrun worker of tests/test_multigpu_gpu.py: one captured training step on this rank's shard of a global batch;
rank 0 saves the all-reduced flat gradient and the updated parameters."""
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    out_dir, graphed = sys.argv[1], sys.argv[2] == "1"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from tests.test_multigpu_gpu import build_case, shard_bounds
    from dgn_b200.engine import TrainStep
    from dgn_b200.graph import collate
    from dgn_b200.parallel import shard_loss_weight
    samples, make_net = build_case(dev)
    lo, hi = shard_bounds(len(samples), world)[rank]
    mine = samples[lo:hi]
    cap = (sum(s["n"] for s in mine) + 40, sum(len(s["src"]) for s in mine) + 64) if graphed else None
    g, labels = collate(mine, capacity=cap)
    tg = labels.float().unsqueeze(1)
    net = make_net()
    step = TrainStep(net, g, tg, lr=1e-3, weight_decay=0.0, graphed=graphed, warmup_iters=2,
                     loss_weight=shard_loss_weight(len(mine), len(samples), world))
    # the capture warm-up already took optimizer steps: restore the initial state, then take ONE measured step
    ref = make_net()
    with torch.no_grad():
        for p, q in zip(net.parameters(), ref.parameters()):
            p.copy_(q)
        for b, c in zip(net.buffers(), ref.buffers()):
            b.copy_(c)
    step.opt.exp_avg.zero_()
    step.opt.exp_avg_sq.zero_()
    step.opt.state.zero_()
    if graphed:
        step.load(collate(mine, capacity=cap)[0], tg.pin_memory())
    else:
        step.g, step.targets = collate(mine)[0].to(dev), tg.to(dev)
    loss = float(step.run())
    torch.cuda.synchronize()
    if rank == 0:
        summed = step.peer.reduced if step.peer is not None else step.flat_g
        torch.save({"peer": None if step.peer is None else ("one-shot" if step.peer.one_shot else "two-shot"),
                    "flat_g": summed.cpu(), "flat_p": step.flat_p.detach().cpu(), "loss0": loss,
                    "launches": step.launches_per_step}, os.path.join(out_dir, "rank0_%d.pt" % int(graphed)))
    assert step.peer is None or not step.peer.timed_out()
    # the replicas must stay bit-identical (deterministic rank-order sum on every rank)
    mine_p = step.flat_p.detach().clone()
    other = [torch.empty_like(mine_p) for _ in range(world)]
    dist.all_gather(other, mine_p)
    assert all(torch.equal(o, other[0]) for o in other), "replicas diverged"
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
