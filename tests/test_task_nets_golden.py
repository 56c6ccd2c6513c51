"""Round-2 task nets (HIV / PCBA with the OGB encoders, towers through the net, virtual node, directional readouts):
the oracle restatement on CPU and the CUDA product path on the GPU, both against golden vectors produced by the
UNMODIFIED reference (oracle/make_golden.py).  Tolerance 1e-5 (north star)."""
import numpy as np
import pytest
import torch

from tests.helpers import assert_close, load_golden, samples_from_golden

MOL_CASES = ["net_hiv_towers", "net_hiv_edge", "net_pcba_vn", "net_pcba_logsum"]
DIR_CASES = ["net_zinc_directional", "net_zinc_directional_abs"]
CLASS_CASES = ["net_sbm_pattern", "net_superpixel_cifar", "net_superpixel_simple_max"]   # BASELINE configs[4] / configs[2] nets


def _mol_params(gold, device):
    p = {k[2:]: gold[k].item() for k in gold if k.startswith("p/")}
    p["avg_d"] = {"log": torch.tensor(float(gold["avg_log"]))}
    p["device"] = device
    return p


def _zinc_params(gold, device):
    ef = bool(gold["edge_feat_flag"])
    return dict(num_atom_type=28, num_bond_type=4, hidden_dim=16, out_dim=16, in_feat_dropout=0.0, dropout=0.0, L=3,
                type_net=str(gold["type_net"]), pos_enc_dim=0, readout=str(gold["readout"]), graph_norm=True,
                batch_norm=True, aggregators=str(gold["aggregators"]), scalers="identity amplification attenuation",
                avg_d={"log": torch.tensor(float(gold["avg_log"]))}, residual=True, edge_feat=ef,
                edge_dim=8 if ef else 0, pretrans_layers=1, posttrans_layers=1, device=device)


def _check(net, scores, loss, gold, rel=1e-5):
    assert_close(scores, gold["scores"], rel, "scores")
    assert_close(loss, gold["loss"], rel, "loss")
    for k, p in net.named_parameters():
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        assert_close(got, gold["grad/" + k], rel, k)


def _load_params(net, gold):
    sd = net.state_dict()
    for k, _ in net.named_parameters():
        sd[k] = torch.from_numpy(gold["sd/" + k])
    net.load_state_dict(sd)


# ---------------------------------------------------------------------------------------------- CPU: oracle pinned
@pytest.mark.parametrize("case", MOL_CASES)
def test_oracle_mol_net_matches_reference(case):
    from oracle.graphs import collate_standin
    from oracle.task_nets import HivNet, PcbaNet
    gold = load_golden(case)
    samples = samples_from_golden(gold)
    g, _, snorm_n, snorm_e = collate_standin(samples)
    kind = str(gold["kind"])
    torch.manual_seed(int(gold["seed"]))
    net = (HivNet if kind == "hiv" else PcbaNet)(_mol_params(gold, "cpu"))
    for k, v in net.state_dict().items():            # same seed + construction order => the reference's initial weights
        if "running" not in k and "num_batches" not in k:
            np.testing.assert_array_equal(v.numpy(), gold["sd/" + k], err_msg=k)
    net.train()
    scores = net(g, g.ndata["feat"], g.edata["feat"], snorm_n, snorm_e)
    loss = net.loss(scores, torch.tensor(gold["targets"]))
    loss.backward()
    _check(net, scores, loss, gold)


@pytest.mark.parametrize("case", DIR_CASES)
def test_oracle_directional_readout_matches_reference(case):
    from oracle.graphs import collate_standin
    from oracle.task_nets import ZincNet
    gold = load_golden(case)
    g, _, snorm_n, snorm_e = collate_standin(samples_from_golden(gold))
    torch.manual_seed(int(gold["seed"]))
    net = ZincNet(_zinc_params(gold, "cpu")).train()
    scores = net(g, g.ndata["feat"], g.edata["feat"], snorm_n, snorm_e)
    loss = net.loss(scores, torch.tensor(gold["targets"]))
    loss.backward()
    _check(net, scores, loss, gold)


@pytest.mark.parametrize("case", CLASS_CASES)
def test_oracle_classification_net_matches_reference(case):
    """SBM node classification (class-balanced cross-entropy, rb/nets/SBMs_node_classification/dgn_net.py:66-81) and
    superpixel graph classification (rb/nets/superpixels_graph_classification/dgn_net.py)."""
    from oracle.graphs import collate_standin
    from oracle.task_nets import PatternNet, SuperpixelNet
    gold = load_golden(case)
    samples = samples_from_golden(gold)
    g, _, snorm_n, snorm_e = collate_standin(samples)
    torch.manual_seed(int(gold["seed"]))
    net = (PatternNet if str(gold["kind"]) == "sbm" else SuperpixelNet)(_mol_params(gold, "cpu")).train()
    _load_params(net, gold)
    scores = net(g, g.ndata["feat"], g.edata["feat"], snorm_n, snorm_e)
    loss = net.loss(scores, torch.tensor(gold["targets"]))
    loss.backward()
    _check(net, scores, loss, gold)


# ---------------------------------------------------------------------------------------------- GPU: product path
@pytest.mark.gpu
@pytest.mark.parametrize("case", CLASS_CASES)
def test_cuda_classification_net_matches_reference(case):
    from dgn_b200.graph import collate
    from dgn_b200.task_nets.SBMs_node_classification import DGNNet as SbmNet
    from dgn_b200.task_nets.superpixels_graph_classification import DGNNet as SpNet
    gold = load_golden(case)
    g, _ = collate(samples_from_golden(gold))
    g.to("cuda")
    net = (SbmNet if str(gold["kind"]) == "sbm" else SpNet)(_mol_params(gold, "cuda"))
    _load_params(net, gold)
    net.to("cuda").train()
    scores = net(g, g.ndata["feat"], g.edata["feat"], g.snorm_n, None)
    loss = net.loss(scores, torch.tensor(gold["targets"], device="cuda"))
    loss.backward()
    _check(net, scores, loss, gold)



@pytest.mark.gpu
@pytest.mark.parametrize("case", MOL_CASES)
def test_cuda_mol_net_matches_reference(case):
    from dgn_b200.graph import collate
    from dgn_b200.task_nets.HIV_graph_classification import DGNNet as HivNet
    from dgn_b200.task_nets.PCBA_graph_classification import DGNNet as PcbaNet
    gold = load_golden(case)
    g, _ = collate(samples_from_golden(gold))
    g.to("cuda")
    kind = str(gold["kind"])
    net = (HivNet if kind == "hiv" else PcbaNet)(_mol_params(gold, "cuda"))
    _load_params(net, gold)
    net.to("cuda").train()
    scores = net(g, g.ndata["feat"], g.edata["feat"], g.snorm_n, None)
    loss = net.loss(scores, torch.tensor(gold["targets"], device="cuda"))
    loss.backward()
    _check(net, scores, loss, gold)


@pytest.mark.gpu
@pytest.mark.parametrize("case", DIR_CASES)
def test_cuda_directional_readout_matches_reference(case):
    from dgn_b200.graph import collate
    from dgn_b200.task_nets.molecules_graph_regression import DGNNet
    gold = load_golden(case)
    g, _ = collate(samples_from_golden(gold))
    g.to("cuda")
    net = DGNNet(_zinc_params(gold, "cuda"))
    _load_params(net, gold)
    net.to("cuda").train()
    scores = net(g, g.ndata["feat"], g.edata["feat"], g.snorm_n, None)
    loss = net.loss(scores, torch.tensor(gold["targets"], device="cuda"))
    loss.backward()
    _check(net, scores, loss, gold)
