"""GPU: device-side collation (dgn_collate_device) is bit-identical to the host collate of the same graphs
(dgl.batch + snorm_n of rb/data/molecules.py:219-230 restated in graph.collate), and a training step fed with an index
list equals the step fed with the host-collated batch."""
import numpy as np
import pytest
import torch

from dgn_b200.data.device_dataset import DeviceDataset
from dgn_b200.data.synthetic import make_samples, avg_log_degree
from dgn_b200.graph import collate

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("kind,kw", [("zinc", {}), ("cifar", dict(n_min=20, n_max=40)), ("molhiv", {}),
                                     ("pattern", dict(n_min=20, n_max=30))])
def test_device_collate_equals_host_collate(kind, kw):
    samples = make_samples(kind, 40, seed=3, **kw)
    graph_level = np.ndim(samples[0]["label"]) == 0
    tg = torch.tensor([[float(s["label"])] for s in samples]) if graph_level else None
    ds = DeviceDataset(samples, DEV, targets=tg)
    rng = np.random.default_rng(0)
    B = 12
    cap = (B * int(ds.sizes.max()), B * int(ds.esizes.max()))     # the trials below repeat graphs
    dev_g = ds.template(B, cap)
    dev_g.bind_device_blob(torch.zeros(dev_g._host_blob.numel(), dtype=torch.uint8, device=DEV))
    tout = torch.zeros(B, 1, device=DEV) if graph_level else None
    for trial in range(4):
        ids = rng.integers(0, len(samples), size=B).astype(np.int32)
        if trial == 3:
            ids[:] = ids[0]                                     # the same graph 12 times
        ds.collate_into(dev_g, torch.from_numpy(ids).to(DEV), tout)
        host_g, _ = collate([samples[i] for i in ids], capacity=cap, graph_capacity=B)
        got = dev_g._blob.cpu().numpy()
        want = host_g._host_blob.numpy()
        if not np.array_equal(got, want):
            hv, dv = host_g._pack.host_views(want), host_g._pack.host_views(got)
            bad = [k for k in hv if not np.array_equal(hv[k], dv[k])]
            raise AssertionError("device collate differs from host collate in %s (trial %d)" % (bad, trial))
        if tout is not None:
            assert torch.equal(tout.cpu(), tg[torch.from_numpy(ids).long()])
    # a batch that does not fit is flagged, not written out of bounds
    small = ds.template(B, (samples[0]["n"] + 8, len(samples[0]["src"]) + 8))
    small.bind_device_blob(torch.zeros(small._host_blob.numel(), dtype=torch.uint8, device=DEV))
    ds.collate_into(small, torch.arange(B, dtype=torch.int32, device=DEV), None)
    assert int(small.meta[3]) == 1
    from dgn_b200._lib import DgnError
    with pytest.raises(DgnError):
        small.check_overflow()
    dev_g.check_overflow()                                       # the batches that fitted are not flagged


def test_train_step_from_index_list_equals_host_batches():
    from dgn_b200.engine import TrainStep
    from dgn_b200.task_nets.molecules_graph_regression import DGNNet
    samples = make_samples("zinc", 64, seed=9)
    avg = avg_log_degree(samples)
    tg = torch.tensor([[float(s["label"])] for s in samples])
    ds = DeviceDataset(samples, DEV, targets=tg)
    B = 16
    cap = ds.capacity_for(B)

    def net():
        p = dict(num_atom_type=28, num_bond_type=4, hidden_dim=32, out_dim=32, in_feat_dropout=0.0, dropout=0.0, L=2,
                 type_net="complex", pos_enc_dim=0, readout="mean", graph_norm=True, batch_norm=True,
                 aggregators="mean max dir1-dx dir2-av", scalers="identity amplification attenuation",
                 avg_d={"log": torch.tensor(avg)}, residual=True, edge_feat=False, edge_dim=0, pretrans_layers=1,
                 posttrans_layers=1, device=DEV)
        torch.manual_seed(41)
        return DGNNet(p).to(DEV).train()

    first = list(range(B))
    a = TrainStep(net(), collate([samples[i] for i in first], capacity=cap, graph_capacity=B)[0], tg[:B], graphed=True,
                  warmup_iters=2)
    b = TrainStep(net(), ds.template(B, cap), tg[:B], graphed=True, warmup_iters=2)
    # the warm-up steps ran on different template contents: re-synchronise the two models before comparing
    b.flat_p.data.copy_(a.flat_p.data)
    for x, y in zip(a.net.buffers(), b.net.buffers()):
        y.copy_(x)
    for o in (a.opt, b.opt):
        o.exp_avg.zero_(); o.exp_avg_sq.zero_(); o.state.zero_()
    rng = np.random.default_rng(1)
    for it in range(5):
        ids = rng.integers(0, len(samples), size=B).astype(np.int32)
        hb, _ = collate([samples[i] for i in ids], capacity=cap, graph_capacity=B)
        a.load(hb, tg[torch.from_numpy(ids).long()].pin_memory())
        la = float(a.run())
        b.load_ids(ds, torch.from_numpy(ids).pin_memory())
        lb = float(b.run())
        assert la == lb, (it, la, lb)
    assert torch.equal(a.flat_p.data, b.flat_p.data)
