"""CPU: the oracle restatement must reproduce the golden vectors the real reference produced."""
import numpy as np
import pytest
import torch

from oracle import mailbox_ops as mo
from oracle.directional_layers import DGNLayer
from oracle.graphs import collate_standin
from oracle.task_nets import ZincNet
from tests.helpers import load_golden, samples_from_golden, state_from_golden, assert_close


def test_registry_names_match_reference():
    gold = load_golden("aggregators")
    assert sorted(mo.AGGREGATORS) == gold["names"].tolist()
    assert len(mo.AGGREGATORS) == 24
    assert sorted(mo.SCALERS) == load_golden("scalers")["names"].tolist()


@pytest.mark.parametrize("name", sorted(mo.AGGREGATORS))
def test_aggregator_bit_exact(name):
    gold = load_golden("aggregators")
    for si in range(len(gold["shapes"])):
        m = torch.tensor(gold["in/%d/msg" % si], requires_grad=True)
        hi = torch.tensor(gold["in/%d/h_in" % si], requires_grad=True)
        y = mo.AGGREGATORS[name](m, torch.tensor(gold["in/%d/eig_s" % si]), torch.tensor(gold["in/%d/eig_d" % si]), hi)
        y.backward(torch.tensor(gold["in/%d/gy" % si]))
        np.testing.assert_array_equal(y.detach().numpy(), gold["out/%d/%s/y" % (si, name)])
        np.testing.assert_array_equal(m.grad.numpy(), gold["out/%d/%s/dmsg" % (si, name)])
        dh = hi.grad.numpy() if hi.grad is not None else np.zeros_like(gold["in/%d/h_in" % si])
        np.testing.assert_array_equal(dh, gold["out/%d/%s/dh_in" % (si, name)])


def test_edge_cases_recorded_in_golden():
    """constant column -> std 1e-4 with zero grad; ties -> first entry; zero field -> 0."""
    gold = load_golden("aggregators")
    si = 2                                                   # shape (5, 3)
    assert np.allclose(gold["out/%d/std/y" % si][1, 2], 1e-4, rtol=1e-3)
    assert np.all(gold["out/%d/std/dmsg" % si][1, :, 2] == 0)
    gmax = gold["out/%d/max/dmsg" % si][0]                   # rows 0 and 1 of node 0 tie
    gy = gold["in/%d/gy" % si][0]
    tied_cols = np.argmax(gold["in/%d/msg" % si][0], axis=0) == 0
    assert np.all(gmax[1][tied_cols] == 0) and np.all(gmax[0][tied_cols] == gy[tied_cols])
    assert np.all(gold["out/%d/dir1-av/y" % si][-1] == 0)


def test_scalers_match():
    gold = load_golden("scalers")
    avg = {"log": torch.tensor(float(gold["avg_log"]), dtype=torch.float32)}
    for D in (1, 2, 3, 4, 9, 51):
        for name in mo.SCALERS:
            got = mo.SCALERS[name](torch.tensor(gold["h"]), D=D, avg_d=avg)
            np.testing.assert_array_equal(np.asarray(got), gold["D%d/%s" % (D, name)])


LAYER_CASES = ["layer_simple", "layer_complex", "layer_complex_extra", "layer_complex_edge",
               "layer_complex_1scaler", "layer_towers", "layer_simple_odd"]


@pytest.mark.parametrize("case", LAYER_CASES)
def test_layer_matches_reference(case):
    gold = load_golden(case)
    g, _, snorm_n, _ = collate_standin(samples_from_golden(gold))
    F, ed = int(gold["F"]), int(gold["edge_dim"])
    layer = DGNLayer(F, F, 0.0, True, True, str(gold["aggregators"]), str(gold["scalers"]),
                     {"log": torch.tensor(float(gold["avg_log"]))}, str(gold["type_net"]), True,
                     towers=int(gold["towers"]), edge_features=ed > 0, edge_dim=ed).model
    layer.load_state_dict(state_from_golden(gold, layer))
    layer.train()
    h = torch.tensor(gold["h"], requires_grad=True)
    e = torch.tensor(gold["e"]) if ed > 0 else None
    y = layer(g, h, e, snorm_n)
    y.backward(torch.tensor(gold["gy"]))
    assert_close(y, gold["y"], 1e-5, "y")
    assert_close(h.grad, gold["dh"], 1e-5, "dh")
    for k, p in layer.named_parameters():
        assert_close(p.grad, gold["grad/" + k], 1e-5, k)
    for k, b in layer.named_buffers():
        if "running" in k:
            assert_close(b, gold["sd/" + k], 1e-5, k)


@pytest.mark.parametrize("case", ["net_zinc_complex", "net_zinc_simple", "net_zinc_edge"])
def test_zinc_net_matches_reference(case):
    gold = load_golden(case)
    samples = samples_from_golden(gold)
    g, _, snorm_n, snorm_e = collate_standin(samples)
    ef = bool(gold["edge_feat_flag"])
    params = dict(num_atom_type=28, num_bond_type=4, hidden_dim=16, out_dim=16, in_feat_dropout=0.0, dropout=0.0,
                  L=3, type_net=str(gold["type_net"]), pos_enc_dim=0, readout="mean", graph_norm=True,
                  batch_norm=True, aggregators=str(gold["aggregators"]),
                  scalers="identity amplification attenuation",
                  avg_d={"log": torch.tensor(float(gold["avg_log"]))}, residual=True, edge_feat=ef,
                  edge_dim=8 if ef else 0, pretrans_layers=1, posttrans_layers=1, device="cpu")
    torch.manual_seed(int(gold["seed"]))
    net = ZincNet(params)
    # same seed + same construction order => same initial weights as the reference, no load needed
    for k, v in net.state_dict().items():
        if "running" not in k and "num_batches" not in k:
            np.testing.assert_array_equal(v.numpy(), gold["sd/" + k], err_msg=k)
    net.train()
    scores = net(g, g.ndata["feat"], g.edata["feat"], snorm_n, snorm_e)
    loss = net.loss(scores, torch.tensor(gold["targets"]))
    loss.backward()
    assert_close(scores, gold["scores"], 1e-5, "scores")
    assert_close(loss, gold["loss"], 1e-5, "loss")
    for k, p in net.named_parameters():
        ref = gold["grad/" + k]
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        assert_close(got, ref, 1e-5, k)
