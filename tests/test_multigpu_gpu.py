"""GPU (>= 2 devices, `gpurun --gpus 2`): the batch-sharded data-parallel step - NCCL all-reduce of the flat gradient and
the Adam update (1 / world folded into the optimizer kernel) - must produce the gradient of the GLOBAL mean loss.

Semantics (declared): BatchNorm statistics are per shard, so the single-GPU comparison evaluates every
shard on its own (its own batch statistics) and adds the gradients weighted by the shard's share of the batch; the shards
are deliberately unequal."""
import os
import signal
import socket
import subprocess
import sys
import tempfile

import pytest
import torch

from tests.helpers import assert_close

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AGGS = "mean max min dir1-dx dir2-dx dir1-dx-no-abs dir1-av"


def build_case(dev):
    from dgn_b200.data.synthetic import make_samples, avg_log_degree
    from dgn_b200.task_nets.molecules_graph_regression import DGNNet
    samples = make_samples("zinc", 26, seed=77)
    avg = avg_log_degree(samples)
    p = dict(num_atom_type=28, num_bond_type=4, hidden_dim=32, out_dim=32, in_feat_dropout=0.0, dropout=0.0, L=3,
             type_net="complex", pos_enc_dim=0, readout="mean", graph_norm=True, batch_norm=True, aggregators=AGGS,
             scalers="identity amplification attenuation", avg_d={"log": torch.tensor(avg)}, residual=True,
             edge_feat=False, edge_dim=0, pretrans_layers=1, posttrans_layers=1, device=dev)

    def make_net():
        torch.manual_seed(41)
        return DGNNet(p).to(dev).train()
    return samples, make_net


def shard_bounds(n, world):
    """Deliberately unequal contiguous shards (first rank gets ~60 %)."""
    if world == 1:
        return [(0, n)]
    first = (n * 3) // 5
    rest = n - first
    cuts = [0, first] + [first + (rest * (r + 1)) // (world - 1) for r in range(world - 1)]
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


@pytest.mark.gpu
@pytest.mark.parametrize("graphed,peer", [(1, "one-shot"), (1, "two-shot"), (0, "one-shot"), (1, None), (0, None)])
def test_two_rank_step_equals_weighted_per_shard_gradients(graphed, peer):
    """peer=1: gradients in NVLink peer memory, dgn_allreduce_adam inside the captured graph; peer=0: NCCL all-reduce +
    dgn_adam_step."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    from dgn_b200.graph import collate
    world = 2
    with tempfile.TemporaryDirectory() as tmp:
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
               "127.0.0.1", "--master-port", str(port), os.path.join(REPO, "tests", "_mgpu_worker.py"), tmp, str(graphed)]
        proc = subprocess.Popen(cmd, cwd=REPO, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                                start_new_session=True, env=dict(os.environ, DGN_PEER_ALLREDUCE=str(int(peer is not None)), DGN_PEER_KEEP_SUM="1",
                                         DGN_AR_ONESHOT_MAX_WORLD="2" if peer == "one-shot" else "0"))
        try:
            out, _ = proc.communicate(timeout=150)
        except subprocess.TimeoutExpired:
            os.killpg(proc.pid, signal.SIGKILL)           # torchrun AND its workers
            out, _ = proc.communicate()
            raise AssertionError("two-rank worker timed out:\n" + out[-3000:])
        assert proc.returncode == 0, out[-4000:]
        got = torch.load(os.path.join(tmp, "rank0_%d.pt" % graphed))
    dev = torch.device("cuda", 0)
    samples, make_net = build_case(dev)
    net = make_net()
    B = len(samples)
    for lo, hi in shard_bounds(B, world):             # every shard with its own BatchNorm statistics
        g, labels = collate(samples[lo:hi])
        g.to(dev)
        scores = net(g, g.ndata["feat"], g.edata["feat"], g.snorm_n, None)
        loss = net.loss(scores, labels.float().unsqueeze(1).to(dev))
        (loss * ((hi - lo) / B)).backward()
    want_g = torch.cat([torch.nn.functional.pad(p.grad.reshape(-1), (0, (-p.numel()) % 4)) for p in net.parameters()])
    assert got["peer"] == peer, "gradient exchange went through %s, expected %s" % (got["peer"], peer)
    assert_close(got["flat_g"] / world, want_g.cpu(), rel=2e-5, what="all-reduced gradient / world")
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    opt.step()
    want_p = torch.cat([torch.nn.functional.pad(p.detach().reshape(-1), (0, (-p.numel()) % 4)) for p in net.parameters()])
    assert_close(got["flat_p"], want_p.cpu(), rel=2e-5, what="parameters after one Adam step")
    assert got["launches"] > 0


def test_shard_loss_weights_reproduce_the_global_mean():
    from dgn_b200.parallel import shard_loss_weight
    sizes, world = [15, 9, 2], 3
    B = sum(sizes)
    w = [shard_loss_weight(n, B, world) for n in sizes]
    # mean over ranks of w_r * (mean over shard r) == mean over the global batch
    vals = [torch.arange(n, dtype=torch.float64) + 3 * i for i, n in enumerate(sizes)]
    global_mean = torch.cat(vals).mean()
    rank_avg = sum(wr * v.mean() for wr, v in zip(w, vals)) / world
    assert abs(float(rank_avg - global_mean)) < 1e-12
