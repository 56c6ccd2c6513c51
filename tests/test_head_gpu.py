"""GPU: the one-launch prediction head (dgn_head_forward / backward) and L1 loss against plain PyTorch fp32 (these are
floating-point kernels: the torch ops they replace are the reference).  Tolerance 1e-5 * max(1, ||ref||_inf)."""
import pytest
import torch

from dgn_b200 import ops
from dgn_b200.nets.mlp_readout_layer import MLPReadout
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _torch_head(head, x):
    for fc in head.FC_layers[:-1]:
        x = torch.relu(fc(x))
    return head.FC_layers[-1](x)


@pytest.mark.parametrize("B,d0,dout", [(128, 64, 1), (77, 64, 1), (1, 64, 1), (300, 64, 10), (40, 36, 3), (1024, 64, 2)])
def test_head_matches_torch(B, d0, dout):
    torch.manual_seed(B + d0)
    head = MLPReadout(d0, dout).to(DEV)
    x = torch.randn(B, d0, device=DEV)
    gy = torch.randn(B, dout, device=DEV)
    assert ops.head_supported(x, head.FC_layers)
    xr = x.clone().requires_grad_(True)
    yr = _torch_head(head, xr)
    yr.backward(gy)
    ref = {k: p.grad.clone() for k, p in head.named_parameters()}
    head.zero_grad(set_to_none=True)
    xm = x.clone().requires_grad_(True)
    ym = head(xm)                                        # fused path
    ym.backward(gy)
    assert_close(ym, yr, what="y")
    assert_close(xm.grad, xr.grad, what="dx")
    for k, p in head.named_parameters():
        assert_close(p.grad, ref[k], what=k)
    # direct accumulation into existing .grad buffers (the engine's mode): a second backward doubles them
    ym2 = head(xm.detach().requires_grad_(True))
    ym2.backward(gy)
    for k, p in head.named_parameters():
        assert_close(p.grad, 2 * ref[k], what="accumulated " + k)
    # determinism
    head.zero_grad(set_to_none=True)
    xa = x.clone().requires_grad_(True)
    head(xa).backward(gy)
    assert torch.equal(xa.grad, xm.grad)


def test_head_falls_back_outside_its_limits():
    head = MLPReadout(64, 1).to(DEV)
    assert not ops.head_supported(torch.randn(5000, 64, device=DEV), head.FC_layers)      # node-level heads (SBM)
    assert not ops.head_supported(torch.randn(8, 64), head.FC_layers)                      # CPU tensor
    y = head(torch.randn(5000, 64, device=DEV))
    assert y.shape == (5000, 1)


@pytest.mark.parametrize("shape", [(128, 1), (7, 3), (1, 1), (1000, 1)])
def test_l1_loss_matches_torch(shape):
    torch.manual_seed(shape[0])
    y = torch.randn(shape, device=DEV)
    t = torch.randn(shape, device=DEV)
    y[0, 0] = t[0, 0]                                    # sign(0) = 0
    yr = y.clone().requires_grad_(True)
    lr = torch.nn.L1Loss()(yr, t)
    (3.0 * lr).backward()
    ym = y.clone().requires_grad_(True)
    lm = ops.l1_loss(ym, t)
    (3.0 * lm).backward()
    assert lm.shape == lr.shape
    assert_close(lm, lr, what="loss")
    assert_close(ym.grad, yr.grad, what="dy")
    assert float(ym.grad[0, 0]) == 0.0
