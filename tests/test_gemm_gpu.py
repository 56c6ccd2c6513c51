"""GPU: the tcgen05 3xTF32 GEMM must be as accurate as an fp32 GEMM (checked against fp64) in every layout."""
import pytest
import torch

from dgn_b200 import _lib, ops
from dgn_b200.ops import gemm


@pytest.fixture(autouse=True)
def _count_tensor_core_launches():
    ops.LAUNCHES = 0
    yield

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _case(M, N, K, a_k, b_k, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    a = torch.randn((M, K) if a_k else (K, M), device=DEV, generator=g)
    b = torch.randn((N, K) if b_k else (K, N), device=DEV, generator=g)
    A = (a if a_k else a.t()).double()
    B = (b.t() if b_k else b).double()
    return a, b, A @ B


SHAPES = [
    (3000, 64, 1984),      # y = cat @ W_post^T          (BASELINE cfg2 posttrans, split-K)
    (3000, 1984, 64),      # d_cat = d_y @ W_post
    (64, 1984, 3000),      # dW_post = d_y^T @ cat       (K = nodes)
    (64, 64, 3000),        # dW_pre blocks
    (3000, 64, 64),        # P = h @ W_src^T
    (128, 64, 32), (129, 68, 36), (1, 4, 4), (257, 130, 100), (2944, 64, 1984),
]


@pytest.mark.parametrize("a_k", [True, False])
@pytest.mark.parametrize("b_k", [True, False])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_matches_fp64(M, N, K, a_k, b_k):
    a, b, ref = _case(M, N, K, a_k, b_k)
    before = _lib.lib.dgn_abi_version()
    out = gemm(a, b, a_kmajor=a_k, b_kmajor=b_k)
    assert before and out.shape == (M, N)
    if K >= 256 and K % 4 == 0 and (M if not a_k else K) % 4 == 0 and (N if not b_k else K) % 4 == 0:
        assert ops.LAUNCHES == 1, "expected the tcgen05 kernel, not the library fallback"
    lib_out = (a if a_k else a.t()) @ (b.t() if b_k else b)                 # fp32 library GEMM
    scale = float(ref.abs().max())
    err = float((out.double() - ref).abs().max())
    err_lib = float((lib_out.double() - ref).abs().max())
    assert err <= max(6 * err_lib, 3e-6 * scale), (err, err_lib, scale)


def test_gemm_accumulate_transposed_and_strided_out():
    a, b, ref = _case(512, 96, 200, False, False, seed=3)                   # dW^T-style product
    base = torch.randn(96, 640, device=DEV)
    out = base.clone()
    view = out[:, 64:576]                                                   # [N=96, M=512] window, row stride 640
    gemm(a, b, a_kmajor=False, b_kmajor=False, out=view, accumulate=True, c_transposed=True)
    exp = base.double()
    exp[:, 64:576] += ref.t()
    assert float((out.double() - exp).abs().max()) <= 2e-6 * float(ref.abs().max()) + 1e-6
    assert torch.equal(out[:, :64], base[:, :64]) and torch.equal(out[:, 576:], base[:, 576:])


def test_gemm_is_deterministic_and_graph_capturable():
    a, b, _ = _case(3000, 64, 1984, True, True, seed=5)
    x = gemm(a, b)
    y = gemm(a, b)
    assert torch.equal(x, y)
    out = torch.empty_like(x)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        gemm(a, b, out=out)
    torch.cuda.current_stream().wait_stream(s)
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg):
        gemm(a, b, out=out)
    out.zero_()
    cg.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, x)


def test_unaligned_shapes_fall_back_to_the_library():
    a, b, ref = _case(50, 45, 1395 // 5 * 5 + 1, True, True)               # K not a multiple of 4
    out = gemm(a, b)
    assert float((out.double() - ref).abs().max()) <= 1e-4 * float(ref.abs().max())


@pytest.mark.parametrize("N,Fi,Fo,extra", [(3000, 64, 64, 0), (257, 16, 16, 8), (33, 20, 20, 0), (1000, 128, 128, 4), (5, 48, 48, 0),
                                           (300, 32, 64, 3), (70, 36, 20, 1)])
def test_pair_linear_matches_fp64(N, Fi, Fo, extra):
    """P = h W_src^T, Q = h W_dst^T and d_h += d_P W_src + d_Q W_dst straight from W = [W_src | W_dst | edge cols]."""
    from dgn_b200.ops import pair_linear_backward, pair_linear_forward
    g = torch.Generator(device=DEV).manual_seed(N)
    h = torch.randn(N, Fi, device=DEV, generator=g)
    W = torch.randn(Fo, 2 * Fi + extra, device=DEV, generator=g)
    P, Q = pair_linear_forward(h, W, Fi)
    assert ops.LAUNCHES == 1
    refP, refQ = h.double() @ W[:, :Fi].double().t(), h.double() @ W[:, Fi:2 * Fi].double().t()
    assert float((P.double() - refP).abs().max()) <= 2e-6 * float(refP.abs().max())
    assert float((Q.double() - refQ).abs().max()) <= 2e-6 * float(refQ.abs().max())
    dP = torch.randn(N, Fo, device=DEV, generator=g)
    dQ = torch.randn(N, Fo, device=DEV, generator=g)
    base = torch.randn(N, Fi, device=DEV, generator=g)
    dh = base.clone()
    pair_linear_backward(dP, dQ, W, Fi, dh)
    ref = base.double() + dP.double() @ W[:, :Fi].double() + dQ.double() @ W[:, Fi:2 * Fi].double()
    assert float((dh.double() - ref).abs().max()) <= 2e-6 * float(ref.abs().max())
