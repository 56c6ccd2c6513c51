"""CPU, build container only: the reference's UNMODIFIED task net picks up this repo's
``nets.dgn_layer`` / ``nets.aggregators`` / ``nets.scalers`` through the namespace-package overlay."""
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/realworld_benchmark"

SCRIPT = r'''
import sys, torch
import nets.dgn_layer, nets.aggregators, nets.scalers, nets.layers
from nets.molecules_graph_regression.dgn_net import DGNNet
import nets.molecules_graph_regression.dgn_net as ref_net
repo, ref = sys.argv[1], sys.argv[2]
assert nets.dgn_layer.__file__.startswith(repo), nets.dgn_layer.__file__
assert nets.aggregators.__file__.startswith(repo) and nets.scalers.__file__.startswith(repo)
assert ref_net.__file__.startswith(ref), ref_net.__file__          # the task net IS the reference file
p = dict(num_atom_type=28, num_bond_type=4, hidden_dim=16, out_dim=16, in_feat_dropout=0.0, dropout=0.0, L=3,
         type_net="complex", pos_enc_dim=0, readout="mean", graph_norm=True, batch_norm=True,
         aggregators="mean dir1-dx dir1-av", scalers="identity amplification attenuation",
         avg_d={"log": torch.tensor(1.1)}, residual=True, edge_feat=False, edge_dim=0, pretrans_layers=1,
         posttrans_layers=1, device="cpu")
torch.manual_seed(41)
net = DGNNet(p)
assert type(net.layers[0]).__module__ == "nets.dgn_layer" and type(net.layers[0]).__name__ == "DGNLayerComplex"
print("KEYS", ",".join(net.state_dict().keys()))
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_reference_task_net_runs_on_the_overlay():
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(
        [REPO, os.path.join(REPO, "dgn_b200"), os.path.join(REPO, "dgn_b200", "compat"), REF]))
    out = subprocess.run([sys.executable, "-c", SCRIPT, REPO, REF], env=env, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-2000:]
    keys = out.stdout.split("KEYS ")[1].strip().split(",")
    import torch
    from oracle.task_nets import ZincNet
    p = dict(num_atom_type=28, num_bond_type=4, hidden_dim=16, out_dim=16, in_feat_dropout=0.0, dropout=0.0, L=3,
             type_net="complex", pos_enc_dim=0, readout="mean", graph_norm=True, batch_norm=True,
             aggregators="mean dir1-dx dir1-av", scalers="identity amplification attenuation",
             avg_d={"log": torch.tensor(1.1)}, residual=True, edge_feat=False, edge_dim=0, pretrans_layers=1,
             posttrans_layers=1, device="cpu")
    assert keys == list(ZincNet(p).state_dict().keys())
