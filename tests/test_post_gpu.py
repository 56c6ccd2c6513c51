"""GPU: the scaler-folded posttrans kernels (dgn_post_forward / backward / wgrad, dgn_pre_wgrad) through the C ABI
against an fp64 evaluation of the reference's expression
    cat = [h | c_0 agg | ... | c_{S-1} agg] ; y = cat W^T            (rb/nets/dgn_layer.py:94-96, :116-119)
Tolerance: 1e-5 of the output scale (north star)."""
import numpy as np
import pytest
import torch

from dgn_b200 import ops
from dgn_b200.graph import BatchedGraph
from dgn_b200.nets.aggregators import AGGREGATORS
from dgn_b200.nets.scalers import SCALERS
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"

# N, F (lead), A, F_agg, F_out, scalers
CASES = [
    (3008, 64, 10, 64, 64, ["identity", "amplification", "attenuation"]),     # bench workload (cfg2)
    (301, 0, 4, 32, 32, ["identity", "amplification", "attenuation"]),        # simple layer: no lead block
    (777, 16, 7, 16, 16, ["amplification"]),                                  # single scaler: not applied
    (1500, 20, 7, 20, 20, ["identity", "attenuation"]),                       # widths off the 32 / 64 grid
    (4500, 48, 5, 48, 96, ["attenuation", "amplification", "identity"]),      # two output tiles, no split-K
    (130, 8, 3, 8, 8, ["identity", "amplification", "attenuation", "identity"]),   # 4 scalers
]


def _setup(case, seed=0):
    N, F, A, Fa, Fo, scalers = case
    rng = np.random.default_rng(seed)
    # a graph whose in-degrees cover 0 (isolated), 1 and larger values
    deg = rng.integers(0, 6, size=N)
    deg[:3] = 0
    dst = np.repeat(np.arange(N), deg).astype(np.int32)
    src = rng.integers(0, N, size=dst.shape[0]).astype(np.int32)
    g = BatchedGraph(N, src, dst).to(DEV)
    aggs = [AGGREGATORS["mean"]] * A
    spec = ops.AggSpec(aggs, [SCALERS[s] for s in scalers], 1.2345, Fa, 3)
    ps = ops.PostSpec(spec, F, Fo)
    assert ps.supported(spec)
    Ka = A * Fa
    cat = torch.tensor(rng.standard_normal((N, F + Ka)).astype(np.float32), device=DEV)
    cat[deg == 0, F:] = 0.0                      # aggregates of isolated nodes are zero rows (DGL 0.4.2)
    W = torch.tensor((rng.standard_normal((Fo, ps.w_cols)) / 8).astype(np.float32), device=DEV)
    ld = np.log(deg + 1.0)
    coefs = []
    S = len(scalers) if len(scalers) > 1 else 0
    for s in scalers[:S]:
        with np.errstate(divide="ignore"):
            c = {"identity": np.ones(N), "amplification": ld / 1.2345, "attenuation": 1.2345 / ld}[s]
        # in-degree 0: the aggregates are zero rows and the kernels use a finite coefficient (1 for identity, 0 else)
        c = np.where(deg == 0, 1.0 if s == "identity" else 0.0, c)
        coefs.append(torch.tensor(c, device=DEV, dtype=torch.float64))
    return g, ps, cat, W, coefs


def _full_cat(cat, F, coefs):
    c64 = cat.double()
    if not coefs:
        return c64
    return torch.cat([c64[:, :F]] + [c64[:, F:] * c[:, None] for c in coefs], dim=1)


@pytest.mark.parametrize("case", range(len(CASES)))
def test_post_forward_backward_wgrad_match_fp64(case):
    N, F, A, Fa, Fo, scalers = CASES[case]
    g, ps, cat, W, coefs = _setup(CASES[case], seed=case)
    full = _full_cat(cat, F, coefs)
    # forward
    y = torch.full((N, Fo), float("nan"), device=DEV)
    ops.post_forward(ps, g, cat, W, y)
    assert_close(y, (full @ W.double().t()).float(), what="y")
    # backward w.r.t. cat
    rng = np.random.default_rng(100 + case)
    d_y = torch.tensor(rng.standard_normal((N, Fo)).astype(np.float32), device=DEV)
    d_cat = torch.full_like(cat, float("nan"))
    ops.post_backward(ps, g, cat, W, d_y, d_cat)
    d_full = d_y.double() @ W.double()
    Ka = A * Fa
    if coefs:
        want = torch.cat([d_full[:, :F]] + [sum(c[:, None] * d_full[:, F + s * Ka:F + (s + 1) * Ka]
                                                for s, c in enumerate(coefs))], dim=1)
    else:
        want = d_full
    assert_close(d_cat, want.float(), what="d_cat")
    # weight gradient, overwrite and accumulate
    want_w = (d_y.double().t() @ full).float()
    d_w = torch.full_like(W, float("nan"))
    ops.post_wgrad(ps, g, cat, W, d_y, d_w, False)
    assert_close(d_w, want_w, what="d_w")
    d_w2 = torch.ones_like(W)
    ops.post_wgrad(ps, g, cat, W, d_y, d_w2, True)
    assert_close(d_w2, want_w + 1.0, what="d_w accumulate")
    # determinism (cluster reduction in rank order)
    y2 = torch.empty_like(y)
    ops.post_forward(ps, g, cat, W, y2)
    d_w3 = torch.empty_like(W)
    ops.post_wgrad(ps, g, cat, W, d_y, d_w3, False)
    assert torch.equal(y, y2) and torch.equal(d_w, d_w3)


@pytest.mark.parametrize("N,Fi,Fo,extra", [(3008, 64, 64, 0), (333, 16, 16, 4), (1000, 20, 20, 0), (70, 48, 96, 8)])
def test_pre_wgrad_matches_fp64(N, Fi, Fo, extra):
    rng = np.random.default_rng(N)
    t = lambda *s: torch.tensor(rng.standard_normal(s).astype(np.float32), device=DEV)
    h, dP, dQ = t(N, Fi), t(N, Fo), t(N, Fo)
    gW = torch.full((Fo, 2 * Fi + extra), 7.0, device=DEV)
    gb = torch.full((Fo,), 3.0, device=DEV)
    assert ops.pre_wgrad(h, dP, dQ, gW, gb, True)
    want = torch.full_like(gW, 7.0)
    want[:, :Fi] += (dP.double().t() @ h.double()).float()
    want[:, Fi:2 * Fi] += (dQ.double().t() @ h.double()).float()
    assert_close(gW, want, what="d_W_pre")
    assert_close(gb, 3.0 + dQ.double().sum(0).float(), what="d_b_pre")
    gW2 = torch.full_like(gW, float("nan"))
    gb2 = torch.full_like(gb, float("nan"))
    assert ops.pre_wgrad(h, dP, dQ, gW2, gb2, False)
    assert_close(gW2[:, :2 * Fi], want[:, :2 * Fi] - 7.0, what="d_W_pre overwrite")
    assert_close(gb2, gb - 3.0, what="d_b_pre overwrite")


def test_folded_layer_equals_concatenated_layer():
    """The whole fused layer with the scalers folded (default) vs the [N, (1+S*A)F] concatenation path (DGN_NO_FOLD)."""
    from dgn_b200.data.synthetic import make_samples, avg_log_degree
    from dgn_b200.graph import collate
    from dgn_b200.nets.dgn_layer import DGNLayer
    samples = make_samples("zinc", 24, seed=5)
    avg = avg_log_degree(samples)
    args = (32, 32, 0.0, True, True, "mean max min std dir1-dx dir2-dx dir1-dx-no-abs dir2-av",
            "identity amplification attenuation", {"log": torch.tensor(avg)}, "complex", True)
    g, _ = collate(samples)
    g.to(DEV)
    torch.manual_seed(3)
    h = torch.randn(g.number_of_nodes(), 32, device=DEV)
    gy = torch.randn(g.number_of_nodes(), 32, device=DEV)
    outs = []
    for fold in (True, False):
        ops.FOLD_ENABLED = fold
        try:
            torch.manual_seed(41)
            layer = DGNLayer(*args, edge_features=False, edge_dim=0).model.to(DEV).train()
            hh = h.clone().requires_grad_(True)
            y = layer(g, hh, None, g.snorm_n)
            y.backward(gy)
            outs.append((y.detach(), hh.grad, {k: p.grad for k, p in layer.named_parameters()}))
        finally:
            ops.FOLD_ENABLED = True
    assert_close(outs[0][0], outs[1][0], what="y")
    assert_close(outs[0][1], outs[1][1], what="d_h")
    for k in outs[0][2]:
        assert_close(outs[0][2][k], outs[1][2][k], rel=2e-5, what=k)
