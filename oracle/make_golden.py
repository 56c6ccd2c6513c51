"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  ``/root/reference`` is a python code drop with
no tests and no fixtures, so the golden vectors are outputs of the reference itself:
``realworld_benchmark/nets/{aggregators,scalers,layers,dgn_layer}.py`` and the ZINC ``DGNNet``
are imported untouched (``sys.path``), on top of the DGL-0.4.2 stand-in in ``oracle/standin``.
The GPU box has no /root/reference; it only reads the committed ``.npz`` files.

    python -m oracle.make_golden            # rewrites tests/golden/
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DGN_REFERENCE", "/root/reference/realworld_benchmark")
OUT = os.path.join(REPO, "tests", "golden")

LAYER_AGGS = "mean max min std dir1-dx dir2-dx dir1-dx-no-abs dir2-dx-no-abs dir1-av dir2-av"
EXTRA_AGGS = "sum var dir3-av dir1-0.1 dir2-neg-0.1 dir3-dx-balanced dir1-dx-balanced"
SCALERS3 = "identity amplification attenuation"


def _import_reference():
    sys.path.insert(0, REPO)
    from oracle import use_standin_dgl
    use_standin_dgl()
    if REF not in sys.path:
        sys.path.insert(1, REF)
    import nets.aggregators as ra
    import nets.scalers as rs
    import nets.dgn_layer as rl
    from nets.molecules_graph_regression.dgn_net import DGNNet
    assert ra.__file__.startswith(REF) and rl.__file__.startswith(REF), "reference modules shadowed"
    return ra, rs, rl, DGNNet


def _np(t):
    return t.detach().cpu().numpy()


def aggregator_cases(ra, rng):
    """Every registry entry on random mailboxes + the edge cases of SURVEY.md 8(c)."""
    out = {}
    F, K = 5, 4
    shapes = [(4, 1), (3, 2), (5, 3), (2, 7)]
    names = sorted(ra.AGGREGATORS)
    out["names"] = np.array(names)
    out["shapes"] = np.array(shapes)
    for si, (n, D) in enumerate(shapes):
        msg = rng.standard_normal((n, D, F)).astype(np.float32)
        eig_s = rng.standard_normal((n, D, K)).astype(np.float32)
        eig_d = np.repeat(rng.standard_normal((n, 1, K)).astype(np.float32), D, axis=1)
        h_in = rng.standard_normal((n, F)).astype(np.float32)
        gy = rng.standard_normal((n, F)).astype(np.float32)
        if D >= 2:
            msg[0, 1] = msg[0, 0]                      # tie in max/min -> gradient to first entry
            msg[1, :, 2] = 0.75                        # constant column -> var 0, std 1e-4
            eig_s[n - 1, :, 1] = eig_d[n - 1, :, 1]    # zero field on eig idx 1 -> weights 0
        out["in/%d/msg" % si], out["in/%d/eig_s" % si] = msg, eig_s
        out["in/%d/eig_d" % si], out["in/%d/h_in" % si], out["in/%d/gy" % si] = eig_d, h_in, gy
        for name in names:
            m = torch.tensor(msg, requires_grad=True)
            hi = torch.tensor(h_in, requires_grad=True)
            y = ra.AGGREGATORS[name](m, torch.tensor(eig_s), torch.tensor(eig_d), hi)
            y.backward(torch.tensor(gy))
            out["out/%d/%s/y" % (si, name)] = _np(y)
            out["out/%d/%s/dmsg" % (si, name)] = _np(m.grad)
            out["out/%d/%s/dh_in" % (si, name)] = _np(hi.grad) if hi.grad is not None else np.zeros_like(h_in)
    return out


def scaler_cases(rs, rng):
    out = {}
    h = rng.standard_normal((3, 7)).astype(np.float32)
    avg = {"log": torch.tensor(1.1348, dtype=torch.float32)}
    out["h"], out["avg_log"] = h, np.float32(avg["log"].item())
    out["names"] = np.array(sorted(rs.SCALERS))
    for D in (1, 2, 3, 4, 9, 51):
        for name in sorted(rs.SCALERS):
            out["D%d/%s" % (D, name)] = _np(rs.SCALERS[name](torch.tensor(h), D=D, avg_d=avg))
    return out


def _flat_state(module, prefix="sd/"):
    return {prefix + k: _np(v) for k, v in module.state_dict().items()}


def layer_case(rl, samples, type_net, F, aggs, scalers, towers, edge_dim, avg_log, seed):
    from oracle.graphs import collate_standin
    g, _, snorm_n, _ = collate_standin(samples)
    torch.manual_seed(seed)
    layer = rl.DGNLayer(in_dim=F, out_dim=F, dropout=0.0, graph_norm=True, batch_norm=True, aggregators=aggs,
                        scalers=scalers, avg_d={"log": torch.tensor(avg_log, dtype=torch.float32)},
                        type_net=type_net, residual=True, towers=towers, divide_input=True,
                        edge_features=edge_dim > 0, edge_dim=edge_dim).model
    layer.train()
    N, E = g.number_of_nodes(), g.number_of_edges()
    h = torch.randn(N, F, requires_grad=True)
    e = torch.randn(E, edge_dim) if edge_dim > 0 else None
    gy = torch.randn(N, F)
    y = layer(g, h, e, snorm_n)
    y.backward(gy)
    out = {"type_net": np.array(type_net), "aggregators": np.array(aggs), "scalers": np.array(scalers),
           "towers": np.int64(towers), "edge_dim": np.int64(edge_dim), "avg_log": np.float32(avg_log),
           "F": np.int64(F), "h": _np(h), "gy": _np(gy), "y": _np(y), "dh": _np(h.grad),
           "snorm_n": _np(snorm_n), "eig": _np(g.ndata["eig"]),
           "src": _np(g.edges()[0]).astype(np.int32), "dst": _np(g.edges()[1]).astype(np.int32),
           "batch_num_nodes": np.array(g.batch_num_nodes, dtype=np.int64)}
    if e is not None:
        out["e"] = _np(e)
    out.update(_flat_state(layer))               # includes BN running stats AFTER the step
    for k, p in layer.named_parameters():
        out["grad/" + k] = _np(p.grad)
    return out


def net_case(DGNNet, samples, avg_log, seed, type_net="complex", aggs="mean dir1-dx", edge_feat=False, readout="mean"):
    from oracle.graphs import collate_standin
    g, labels, snorm_n, snorm_e = collate_standin(samples)
    params = dict(num_atom_type=28, num_bond_type=4, hidden_dim=16, out_dim=16, in_feat_dropout=0.0, dropout=0.0,
                  L=3, type_net=type_net, pos_enc_dim=0, readout=readout, graph_norm=True, batch_norm=True,
                  aggregators=aggs, scalers=SCALERS3, avg_d={"log": torch.tensor(avg_log, dtype=torch.float32)},
                  residual=True, edge_feat=edge_feat, edge_dim=8 if edge_feat else 0, pretrans_layers=1,
                  posttrans_layers=1, device="cpu")
    torch.manual_seed(seed)
    net = DGNNet(params)
    net.train()
    x, e = g.ndata["feat"], g.edata["feat"]
    scores = net.forward(g, x, e, snorm_n, snorm_e)
    targets = labels.float().unsqueeze(1)
    loss = net.loss(scores, targets)
    loss.backward()
    out = {"node_feat": _np(x), "edge_feat": _np(e), "targets": _np(targets), "scores": _np(scores),
           "loss": _np(loss), "snorm_n": _np(snorm_n), "eig": _np(g.ndata["eig"]), "avg_log": np.float32(avg_log),
           "src": _np(g.edges()[0]).astype(np.int32), "dst": _np(g.edges()[1]).astype(np.int32),
           "batch_num_nodes": np.array(g.batch_num_nodes, dtype=np.int64), "type_net": np.array(type_net),
           "aggregators": np.array(aggs), "edge_feat_flag": np.int64(edge_feat), "seed": np.int64(seed),
           "readout": np.array(readout)}
    out.update(_flat_state(net))
    for k, p in net.named_parameters():
        out["grad/" + k] = _np(p.grad) if p.grad is not None else np.zeros(tuple(p.shape), np.float32)
    return out


def mol_net_case(kind, samples, avg_log, seed, **over):
    """HIV / PCBA DGNNet of the unmodified reference (on the ogb stand-in) - scores, loss, all parameter gradients."""
    from oracle.graphs import collate_standin
    if kind == "hiv":
        from nets.HIV_graph_classification.dgn_net import DGNNet
    else:
        from nets.PCBA_graph_classification.dgn_net import DGNNet
    g, labels, snorm_n, snorm_e = collate_standin(samples)
    params = dict(hidden_dim=20, out_dim=20, in_feat_dropout=0.0, dropout=0.0, L=3, type_net="towers", pos_enc_dim=0,
                  readout="mean", graph_norm=True, batch_norm=True,
                  aggregators="mean max min dir1-dx dir2-dx dir1-av dir2-av", scalers="identity",
                  avg_d={"log": torch.tensor(avg_log, dtype=torch.float32)}, residual=True, edge_feat=False, edge_dim=0,
                  pretrans_layers=1, posttrans_layers=1, device="cpu", towers=5, decreasing_dim=True,
                  virtual_node="none")
    params.update(over)
    torch.manual_seed(seed)
    net = DGNNet(params)
    net.train()
    x, e = g.ndata["feat"], g.edata["feat"]
    scores = net.forward(g, x, e, snorm_n, snorm_e)
    if kind == "hiv":          # the reference's loss moves the labels to 'cuda' unconditionally (dgn_net.py:88)
        targets = labels.float()
        loss = torch.nn.BCEWithLogitsLoss()(scores, targets.unsqueeze(-1))
    else:
        rng = np.random.default_rng(seed)
        targets = torch.tensor((rng.random((len(samples), 128)) < 0.3).astype(np.float32))
        loss = net.loss(scores, targets)
    loss.backward()
    out = {"node_feat": _np(x), "edge_feat": _np(e), "targets": _np(targets), "scores": _np(scores), "loss": _np(loss),
           "snorm_n": _np(snorm_n), "eig": _np(g.ndata["eig"]), "avg_log": np.float32(avg_log),
           "src": _np(g.edges()[0]).astype(np.int32), "dst": _np(g.edges()[1]).astype(np.int32),
           "batch_num_nodes": np.array(g.batch_num_nodes, dtype=np.int64), "seed": np.int64(seed),
           "kind": np.array(kind)}
    for k, v in params.items():
        if isinstance(v, (int, float, bool, str)):
            out["p/" + k] = np.array(v)
    out.update(_flat_state(net))
    for k, p in net.named_parameters():
        out["grad/" + k] = _np(p.grad) if p.grad is not None else np.zeros(tuple(p.shape), np.float32)
    return out


def class_net_case(kind, samples, avg_log, seed, **over):
    """SBM (node classification, class-balanced CE) / superpixel (graph classification) DGNNet of the unmodified
    reference - scores, loss, all parameter gradients."""
    from oracle.graphs import collate_standin
    if kind == "sbm":
        from nets.SBMs_node_classification.dgn_net import DGNNet
    else:
        from nets.superpixels_graph_classification.dgn_net import DGNNet
    g, labels, snorm_n, snorm_e = collate_standin(samples)
    x, e = g.ndata["feat"], g.edata["feat"]
    params = dict(hidden_dim=16, out_dim=16, in_feat_dropout=0.0, dropout=0.0, L=3, type_net="complex", pos_enc_dim=0,
                  readout="mean", graph_norm=True, batch_norm=True, aggregators="mean dir1-dx dir2-dx",
                  scalers=SCALERS3, avg_d={"log": torch.tensor(avg_log, dtype=torch.float32)}, residual=True,
                  edge_feat=False, edge_dim=0, pretrans_layers=1, posttrans_layers=1, device="cpu")
    if kind == "sbm":
        params.update(in_dim=3, n_classes=2)
    else:
        params.update(in_dim=int(x.shape[1]), in_dim_edge=1, n_classes=10)
    params.update(over)
    torch.manual_seed(seed)
    net = DGNNet(params)
    net.train()
    scores = net.forward(g, x, e, snorm_n, snorm_e)
    targets = labels.long()
    loss = net.loss(scores, targets)
    loss.backward()
    out = {"node_feat": _np(x), "edge_feat": _np(e), "targets": _np(targets), "scores": _np(scores), "loss": _np(loss),
           "snorm_n": _np(snorm_n), "eig": _np(g.ndata["eig"]), "avg_log": np.float32(avg_log),
           "src": _np(g.edges()[0]).astype(np.int32), "dst": _np(g.edges()[1]).astype(np.int32),
           "batch_num_nodes": np.array(g.batch_num_nodes, dtype=np.int64), "seed": np.int64(seed),
           "kind": np.array(kind)}
    for k, v in params.items():
        if isinstance(v, (int, float, bool, str)):
            out["p/" + k] = np.array(v)
    out.update(_flat_state(net))
    for k, p in net.named_parameters():
        out["grad/" + k] = _np(p.grad) if p.grad is not None else np.zeros(tuple(p.shape), np.float32)
    return out


def main(only_new=False):
    sys.path.insert(0, REPO)
    from dgn_b200.data.synthetic import make_samples, avg_log_degree
    ra, rs, rl, DGNNet = _import_reference()
    torch.set_num_threads(1)
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(2020)

    if not only_new:
        np.savez_compressed(os.path.join(OUT, "aggregators.npz"), **aggregator_cases(ra, rng))
        np.savez_compressed(os.path.join(OUT, "scalers.npz"), **scaler_cases(rs, rng))

    zinc = make_samples("zinc", 6, seed=7)
    cifar = make_samples("cifar", 2, seed=8, n_min=20, n_max=30)     # directed kNN, in-degree 0 possible
    for s in cifar:                                                   # widen eig to 4 columns for dir3-*
        s["eig"] = np.concatenate([s["eig"], s["eig"][:, 1:2] * s["eig"][:, 2:3]], axis=1)
    avg_z, avg_c = avg_log_degree(zinc), avg_log_degree(cifar)
    cases = {
        "layer_simple": layer_case(rl, zinc, "simple", 16, LAYER_AGGS, SCALERS3, 1, 0, avg_z, 11),
        "layer_complex": layer_case(rl, zinc, "complex", 16, LAYER_AGGS, SCALERS3, 1, 0, avg_z, 12),
        "layer_complex_extra": layer_case(rl, cifar, "complex", 12, EXTRA_AGGS, SCALERS3, 1, 0, avg_c, 13),
        "layer_complex_edge": layer_case(rl, zinc, "complex", 16, "mean dir1-dx dir1-av", SCALERS3, 1, 6, avg_z, 14),
        "layer_complex_1scaler": layer_case(rl, cifar, "complex", 13, "mean dir1-dx dir2-dx", "amplification", 1, 0,
                                            avg_c, 15),
        "layer_towers": layer_case(rl, zinc, "towers", 16, "mean max min dir1-dx dir2-dx dir1-av dir2-av",
                                   "identity", 4, 0, avg_z, 16),
        "layer_simple_odd": layer_case(rl, cifar, "simple", 7, "mean std dir1-dx dir2-av max", SCALERS3, 1, 0,
                                       avg_c, 17),
        "net_zinc_complex": net_case(DGNNet, zinc, avg_z, 41, "complex", "mean dir1-dx"),
        "net_zinc_simple": net_case(DGNNet, zinc, avg_z, 41, "simple", "mean max dir1-dx dir1-av"),
        "net_zinc_edge": net_case(DGNNet, zinc, avg_z, 41, "complex", "mean dir1-dx dir2-av", edge_feat=True),
    }
    if only_new:
        cases = {}
    # round 2: the remaining task nets (OGB encoders, towers through the net, virtual node) and the directional readouts
    hiv = make_samples("molhiv", 5, seed=21)
    avg_h = avg_log_degree(hiv)
    cases.update({
        "net_hiv_towers": mol_net_case("hiv", hiv, avg_h, 41),                          # 5 towers (the layer default)
        "net_hiv_edge": mol_net_case("hiv", hiv, avg_h, 42, type_net="complex", edge_feat=True, edge_dim=8,
                                     aggregators="mean dir1-dx dir1-av", scalers=SCALERS3),
        "net_pcba_vn": mol_net_case("pcba", hiv, avg_h, 43, towers=4, virtual_node="mean"),
        "net_pcba_logsum": mol_net_case("pcba", hiv, avg_h, 44, towers=2, virtual_node="logsum", decreasing_dim=False,
                                        residual=False, hidden_dim=16, out_dim=16),
        "net_zinc_directional": net_case(DGNNet, zinc, avg_z, 45, "simple", "mean dir1-dx dir1-av", readout="directional"),
        "net_zinc_directional_abs": net_case(DGNNet, zinc, avg_z, 46, "complex", "mean max dir2-dx",
                                             readout="directional_abs"),
    })
    # the node-classification (SBM PATTERN, BASELINE configs[4]) and superpixel (CIFAR10, configs[2]) nets
    pat = make_samples("pattern", 3, seed=31, n_min=20, n_max=30)
    cif = make_samples("cifar", 4, seed=32, n_min=20, n_max=30)
    cases.update({
        "net_sbm_pattern": class_net_case("sbm", pat, avg_log_degree(pat), 47, aggregators="mean dir1-dx dir2-dx dir3-dx"),
        "net_superpixel_cifar": class_net_case("superpixel", cif, avg_log_degree(cif), 48, hidden_dim=20, out_dim=20,
                                               scalers="identity", readout="mean"),
        "net_superpixel_simple_max": class_net_case("superpixel", cif, avg_log_degree(cif), 49, type_net="simple",
                                                    aggregators="mean max dir1-av", readout="max"),
    })
    for name, payload in cases.items():
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **payload)
    for f in sorted(os.listdir(OUT)):
        print("%-32s %8d bytes" % (f, os.path.getsize(os.path.join(OUT, f))))


if __name__ == "__main__":
    main(only_new="--only-new" in sys.argv)
