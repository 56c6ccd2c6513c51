"""CPU (gloo, world_size 2): the data-parallel plumbing - graph sharding, flat parameter/gradient
buffers and the single gradient all-reduce - without touching the CUDA path."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dgn_b200.data.synthetic import make_samples
from dgn_b200.parallel import allreduce_mean_, flatten_parameters, shard_samples


def test_shard_samples_partitions_and_balances():
    samples = make_samples("zinc", 64, seed=0)
    for world in (1, 2, 4, 8):
        shards = [shard_samples(samples, r, world) for r in range(world)]
        assert sum(len(s) for s in shards) == 64 and all(len(s) > 0 for s in shards)
        flat = [id(x) for s in shards for x in s]
        assert flat == [id(x) for x in samples]                       # contiguous, order preserving
        edges = np.array([sum(len(x["src"]) for x in s) for s in shards], dtype=float)
        assert edges.max() <= 1.35 * edges.mean()


def test_flatten_parameters_keeps_views_live():
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.BatchNorm1d(7), torch.nn.Linear(7, 3))
    ref = [p.detach().clone() for p in m.parameters()]
    flat_p, flat_g = flatten_parameters(m)
    assert all(torch.equal(a, b) for a, b in zip(ref, m.parameters()))
    x = torch.randn(11, 5)
    m(x).sum().backward()
    assert flat_g.abs().sum() > 0 and flat_p.grad is flat_g
    g0 = [p.grad.clone() for p in m.parameters()]
    opt = torch.optim.SGD([flat_p], lr=0.5)
    opt.step()
    for p, r, g in zip(m.parameters(), ref, g0):
        assert torch.allclose(p, r - 0.5 * g)                         # module parameters moved with the flat buffer
    flat_g.zero_()
    assert all(float(p.grad.abs().sum()) == 0 for p in m.parameters())


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(7)
    m = torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.ReLU(), torch.nn.Linear(4, 1))
    flat_p, flat_g = flatten_parameters(m)
    data = torch.randn(16, 6, generator=torch.Generator().manual_seed(1))
    tgt = torch.randn(16, 1, generator=torch.Generator().manual_seed(2))
    sl = slice(rank * 8, (rank + 1) * 8)
    torch.nn.functional.l1_loss(m(data[sl]), tgt[sl]).backward()
    allreduce_mean_(flat_g)
    out[rank] = flat_g.clone()
    dist.destroy_process_group()


def test_two_rank_gradient_equals_global_batch_gradient():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    torch.manual_seed(7)
    m = torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.ReLU(), torch.nn.Linear(4, 1))
    flat_p, flat_g = flatten_parameters(m)
    data = torch.randn(16, 6, generator=torch.Generator().manual_seed(1))
    tgt = torch.randn(16, 1, generator=torch.Generator().manual_seed(2))
    torch.nn.functional.l1_loss(m(data), tgt).backward()              # mean over the global batch
    assert torch.allclose(out[0], out[1])
    assert torch.allclose(out[0], flat_g, atol=1e-6)
