"""torch.autograd wrappers around the C-ABI kernels (``include/dgn_b200.h``).

PyTorch is plumbing here: it owns device memory and the CUDA stream, and records the autograd
graph.  All arithmetic of the hot path happens inside ``libdgn_b200.so``.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib
from ._lib import lib, check


LAUNCHES = 0        # kernels of libdgn_b200.so launched so far (bench.py reports it as gpu_launches)


def _count(n: int) -> None:
    global LAUNCHES
    LAUNCHES += n


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.DgnError("dgn_b200 ops run on CUDA tensors only (got a %s tensor); there is no CPU fallback"
                                % t.device)


def _f32c(t):
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.stride(-1) == 1 and t.dim() == 2 else t.contiguous()


class AggSpec:
    """Aggregator / scaler selection of one layer, frozen into a ``DgnAggSpec``."""

    def __init__(self, aggregators, scalers, avg_log: float, n_feat: int, n_eig: int, group_feat: int = 0):
        if len(aggregators) > _lib.MAX_AGG or len(scalers) > _lib.MAX_SCALERS:
            raise _lib.DgnError("at most %d aggregators and %d scalers per layer" % (_lib.MAX_AGG, _lib.MAX_SCALERS))
        self.aggregators, self.scalers = list(aggregators), list(scalers)
        self.A, self.S_decl = len(self.aggregators), len(self.scalers)
        self.S = self.S_decl if self.S_decl > 1 else 1          # rb/nets/dgn_layer.py:95
        self.F, self.K = int(n_feat), int(n_eig)
        self.Fg = int(group_feat) if group_feat else self.F
        self.avg_log = float(avg_log)
        s = _lib.DgnAggSpec()
        s.n_feat, s.group_feat, s.n_eig, s.n_agg, s.n_scalers = self.F, self.Fg, self.K, self.A, self.S_decl
        for i, a in enumerate(self.aggregators):
            s.agg_kind[i], s.agg_eig[i], s.agg_alpha[i] = a.kind, a.eig_idx, a.alpha
            if a.kind >= _lib.AGG_DIR_AV and a.eig_idx >= self.K:
                raise _lib.DgnError("aggregator %r needs eigenvector column %d but ndata['eig'] has %d columns"
                                    % (a.name, a.eig_idx, self.K))
        for i, sc in enumerate(self.scalers):
            s.scaler_kind[i] = sc.kind
        s.avg_log = self.avg_log
        self.c = s

    @property
    def out_width(self) -> int:
        """Columns of the aggregate per tower: S*A*F_tower."""
        return self.S * self.A * self.Fg


# False (or DGN_NO_FIELD=1): the kernels derive the eigen-weights from eig inside every launch (the ABI v2 kernels)
FIELD_ENABLED = os.environ.get("DGN_NO_FIELD", "0") != "1"


def _agg_io(mode, x, q, r, h_in, eig, out, lead, Wt, cat_input, q_bias=None, field=None):
    io = _lib.DgnAggIO()
    io.msg_mode = mode
    if field is not None:
        io.field = C.addressof(field)
    if q_bias is not None:
        io.q_bias = q_bias.data_ptr()
    if x is not None:
        io.x, io.ld_x = x.data_ptr(), x.stride(0)
    if q is not None:
        io.q, io.ld_q = q.data_ptr(), q.stride(0)
    if r is not None:
        io.r, io.ld_r = r.data_ptr(), r.stride(0)
    io.h_in, io.ld_h = h_in.data_ptr(), h_in.stride(0)
    if eig is not None:
        io.eig, io.ld_eig = eig.data_ptr(), eig.stride(0)
    io.out, io.ld_out, io.out_group_stride = out.data_ptr() + 4 * lead, out.stride(0), Wt
    io.ld_hcopy, io.hcopy_group_stride = out.stride(0), Wt
    if cat_input:
        io.h_copy = out.data_ptr()
    return io


def agg_forward_raw(graph, spec, mode, x, q, r, h_in, eig, out, cat_input, q_bias=None):
    """dgn_agg_forward on pre-allocated tensors (fp32, unit inner stride).  ``out`` is
    ``[N, T*((F_t if cat_input else 0) + S*A*F_t)]``."""
    if graph.number_of_nodes() == 0:
        return                                   # empty batch: nothing to launch (zero-size tensors have no address)
    lead = spec.Fg if cat_input else 0
    Wt = lead + spec.out_width
    field = graph.field(spec, eig) if FIELD_ENABLED else None
    io = _agg_io(mode, x, q, r, h_in, eig, out, lead, Wt, cat_input, q_bias, field)
    check(lib.dgn_agg_forward(C.byref(graph.c_graph()), C.byref(spec.c), C.byref(io), _stream(h_in)),
          "dgn_agg_forward")
    _count(1)


def agg_backward_raw(graph, spec, mode, x, q, r, h_in, eig, g_out, cat_input, d_x=None, d_q=None, d_r=None,
                     d_h=None, edge_ws=None, fold_h_in=False, q_bias=None, d_h_addend=None):
    """dgn_agg_backward on pre-allocated tensors; any of the d_* outputs may be None."""
    if graph.number_of_nodes() == 0:
        return
    lead = spec.Fg if cat_input else 0
    Wt = lead + spec.out_width
    field = graph.field(spec, eig) if FIELD_ENABLED else None
    io = _agg_io(mode, x, q, r, h_in, eig, g_out, lead, Wt, False, q_bias, field)     # io.out is never written here
    gr = _lib.DgnAggGrad()
    gr.g_out = g_out.data_ptr() + 4 * lead
    if cat_input:
        gr.g_hcopy = g_out.data_ptr()
    if d_x is not None:
        gr.d_x, gr.ld_dx, gr.edge_ws = d_x.data_ptr(), d_x.stride(0), edge_ws.data_ptr()
    elif edge_ws is not None:
        gr.edge_ws = edge_ws.data_ptr()          # spill only: the caller reduces over the out-edges (pair_gather_backward)
    if d_q is not None:
        gr.d_q, gr.ld_dq = d_q.data_ptr(), d_q.stride(0)
    if d_r is not None:
        gr.d_r, gr.ld_dr = d_r.data_ptr(), d_r.stride(0)
    if d_h is not None:
        gr.d_h_in, gr.ld_dh = d_h.data_ptr(), d_h.stride(0)
        if d_h_addend is not None:
            gr.d_h_addend, gr.ld_dha = d_h_addend.data_ptr(), d_h_addend.stride(0)
    gr.fold_h_in = 1 if fold_h_in else 0
    check(lib.dgn_agg_backward(C.byref(graph.c_graph()), C.byref(spec.c), C.byref(io), C.byref(gr), _stream(h_in)),
          "dgn_agg_backward")
    _count(2 if d_x is not None else 1)


class _Aggregate(torch.autograd.Function):
    """out = [h_in |] scalers(aggregators(messages)); see dgn_agg_forward / dgn_agg_backward."""

    @staticmethod
    def forward(ctx, graph, spec, mode, cat_input, x, q, r, h_in, eig):
        _need_cuda(x, q, r, h_in, eig)
        x, q, r, h_in, eig = _f32c(x), _f32c(q), _f32c(r), _f32c(h_in), _f32c(eig)
        N, T = graph.number_of_nodes(), spec.F // spec.Fg
        Wt = (spec.Fg if cat_input else 0) + spec.out_width            # columns per tower block
        out = torch.empty((N, T * Wt), device=h_in.device, dtype=torch.float32)
        agg_forward_raw(graph, spec, mode, x, q, r, h_in, eig, out, cat_input)
        ctx.graph, ctx.spec, ctx.mode, ctx.cat_input = graph, spec, mode, cat_input
        ctx.save_for_backward(x, q, r, h_in, eig)
        ctx.same_x_h = x is not None and h_in.data_ptr() == x.data_ptr() and mode == _lib.MSG_SOURCE
        return out

    @staticmethod
    def backward(ctx, g_out):
        x, q, r, h_in, eig = ctx.saved_tensors
        graph, spec, mode = ctx.graph, ctx.spec, ctx.mode
        g_out = g_out.contiguous()
        N, E, F = graph.number_of_nodes(), graph.number_of_edges(), spec.F
        dev = h_in.device
        need = ctx.needs_input_grad        # (graph, spec, mode, cat_input, x, q, r, h_in, eig)
        fold = ctx.same_x_h                # simple layer: x and h_in are one tensor
        d_x = d_q = d_r = d_h = ws = None
        if x is not None and (need[4] or (fold and need[7])):
            d_x = torch.empty((N, F), device=dev, dtype=torch.float32)
            ws = torch.empty((max(E, 1), F), device=dev, dtype=torch.float32)
        if q is not None and need[5]:
            d_q = torch.empty((N, F), device=dev, dtype=torch.float32)
        if r is not None and need[6]:
            # zeros: the kernels write the rows of REAL edges only; padding rows of a fixed-capacity batch must not
            # carry allocator garbage into the caller's weight gradient (d_r^T @ e)
            d_r = torch.zeros((max(E, 1), F), device=dev, dtype=torch.float32)[:E]
        if need[7] or fold:
            d_h = torch.empty((N, F), device=dev, dtype=torch.float32)
        agg_backward_raw(graph, spec, mode, x, q, r, h_in, eig, g_out, ctx.cat_input, d_x, d_q, d_r, d_h, ws,
                         fold_h_in=fold and d_x is not None)
        if fold:
            # x and h_in are the same tensor: its whole gradient was folded into d_x (None counts as zero)
            return None, None, None, None, d_x, None, None, None, None
        return None, None, None, None, d_x, d_q, d_r, d_h, None


def aggregate(graph, spec: AggSpec, mode: int, h_in, eig, x=None, q=None, r=None, cat_input=False):
    """Fused directional aggregation.  Returns ``[N, T*((1 if cat_input else 0)*F_t + S*A*F_t)]``."""
    return _Aggregate.apply(graph, spec, mode, cat_input, x, q, r, h_in, eig)


def norm_forward_raw(y, out, stats, snorm=None, y_bias=None, gamma=None, beta=None, running_mean=None,
                     running_var=None, momentum=0.1, eps=1e-5, training=True, relu=True, residual=None,
                     n_rows_dev=None, stat_parts=0, launch=True):
    """dgn_norm_forward on pre-allocated tensors; returns the filled DgnNormArgs (needed by the backward)."""
    N, Cn = y.shape
    a = _lib.DgnNormArgs()
    a.n_rows, a.n_cols, a.y, a.ld_y = N, Cn, y.data_ptr(), y.stride(0)
    if y_bias is not None:
        a.y_bias = y_bias.data_ptr()
    if snorm is not None:
        a.snorm = snorm.data_ptr()
    if gamma is not None:
        a.gamma, a.beta = gamma.data_ptr(), beta.data_ptr()
        if running_mean is not None:
            a.running_mean, a.running_var = running_mean.data_ptr(), running_var.data_ptr()
    a.momentum, a.eps, a.training, a.relu = momentum, eps, int(training), int(relu)
    if residual is not None:
        a.residual, a.ld_res = residual.data_ptr(), residual.stride(0)
    a.out, a.ld_o, a.stats = out.data_ptr(), out.stride(0), stats.data_ptr()
    if n_rows_dev is not None:
        a.n_rows_dev = n_rows_dev.data_ptr()
    a.stat_parts = int(stat_parts)
    if launch:
        check(lib.dgn_norm_forward(C.byref(a), _stream(y)), "dgn_norm_forward")
        _count(2 if (gamma is not None and training and not stat_parts) else 1)
    return a


def norm_pair_forward(a, W_next, P, Q):
    """The apply pass of ``a`` (a filled DgnNormArgs whose statistics are final) fused with ``P = out W_src^T``,
    ``Q = out W_dst^T`` of the next layer's pretrans weight; False when the shapes are outside the kernel's range."""
    rc = lib.dgn_norm_pair_forward(C.byref(a), P.shape[1], W_next.data_ptr(), W_next.stride(0), P.data_ptr(), P.stride(0),
                                   Q.data_ptr(), Q.stride(0), torch.cuda.current_stream(P.device).cuda_stream)
    if rc == -2:
        return False
    check(rc, "dgn_norm_pair_forward")
    _count(1)
    return True


def norm_backward_raw(a, g_out, d_y, scratch, d_gamma=None, d_beta=None, d_bias=None, accumulate=False):
    g = _lib.DgnNormGrad()
    g.g_out, g.ld_go, g.d_y, g.ld_dy, g.scratch = (g_out.data_ptr(), g_out.stride(0), d_y.data_ptr(), d_y.stride(0),
                                                  scratch.data_ptr())
    if d_gamma is not None:
        g.d_gamma, g.d_beta = d_gamma.data_ptr(), d_beta.data_ptr()
    if d_bias is not None:
        g.d_bias = d_bias.data_ptr()
    g.accumulate = int(accumulate)
    check(lib.dgn_norm_backward(C.byref(a), C.byref(g), torch.cuda.current_stream(d_y.device).cuda_stream),
          "dgn_norm_backward")
    _count(2)


class _NormAct(torch.autograd.Function):
    """snorm * y -> BatchNorm1d -> ReLU -> + residual (rb/nets/dgn_layer.py:122-130) in the C-ABI kernels."""

    @staticmethod
    def forward(ctx, y, snorm, gamma, beta, running_mean, running_var, momentum, eps, training, relu, residual,
                n_rows_dev=None):
        _need_cuda(y)
        y = _f32c(y)
        N, Cn = y.shape
        out = torch.empty((N, Cn), device=y.device, dtype=torch.float32)
        stats = torch.empty(_lib.NORM_WS_PER_COL * Cn, device=y.device, dtype=torch.float32)
        if snorm is not None:
            snorm = snorm.reshape(-1).contiguous()
        if residual is not None:
            residual = _f32c(residual)
        a = norm_forward_raw(y, out, stats, snorm, None, gamma, beta, running_mean, running_var, momentum, eps,
                             training, relu, residual, n_rows_dev)
        ctx.args = a
        ctx.keep = (y, snorm, gamma, beta, running_mean, running_var, residual, out, stats, n_rows_dev)
        ctx.has_res, ctx.has_bn = residual is not None, gamma is not None
        return out

    @staticmethod
    def backward(ctx, g_out):
        y = ctx.keep[0]
        N, Cn = y.shape
        g_out = g_out.contiguous()
        d_y = torch.empty_like(y)
        scratch = torch.empty(_lib.NORM_WS_PER_COL * Cn, device=y.device, dtype=torch.float32)
        d_gamma = d_beta = None
        if ctx.has_bn:
            d_gamma = torch.empty(Cn, device=y.device, dtype=torch.float32)
            d_beta = torch.empty(Cn, device=y.device, dtype=torch.float32)
        norm_backward_raw(ctx.args, g_out, d_y, scratch, d_gamma, d_beta)
        d_res = g_out if ctx.has_res else None       # identity branch
        return d_y, None, d_gamma, d_beta, None, None, None, None, None, None, d_res, None


# nn.BatchNorm1d counts its training batches in `num_batches_tracked`.  Incrementing it costs one tiny launch per
# layer on the step's critical path; a step runner may collect the counters here instead (set BN_COUNTERS to a list)
# and bump them all with ONE multi-tensor launch at the end of the forward pass (engine.TrainStep does).
BN_COUNTERS = None


def count_bn_batch(bn):
    if bn.track_running_stats and bn.num_batches_tracked is not None:
        if BN_COUNTERS is None:
            bn.num_batches_tracked += 1
        else:
            BN_COUNTERS.append(bn.num_batches_tracked)


def norm_act(y, snorm, bn, training, relu, residual, n_rows_dev=None):
    """Fused layer epilogue.  ``bn`` is an ``nn.BatchNorm1d`` (or None when batch_norm is off).

    ``n_rows_dev`` (device int32, optional) is the number of REAL rows of a padded batch: statistics
    run over those rows only and the padding rows of the output are written as zeros."""
    if bn is not None:
        if training:
            count_bn_batch(bn)
        use_batch = training or not bn.track_running_stats
        mom = 0.1 if bn.momentum is None else bn.momentum
        return _NormAct.apply(y, snorm, bn.weight, bn.bias, bn.running_mean, bn.running_var, mom, bn.eps,
                              use_batch, relu, residual, n_rows_dev)
    return _NormAct.apply(y, snorm, None, None, None, None, 0.1, 1e-5, False, relu, residual, n_rows_dev)


class _Readout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, graph, op):
        _need_cuda(h)
        h = _f32c(h)
        B, Cn = graph.batch_size, h.shape[1]
        out = torch.empty((B, Cn), device=h.device, dtype=torch.float32)
        check(lib.dgn_readout_forward(B, graph.graph_ptr.data_ptr(), Cn, h.data_ptr(), h.stride(0), op,
                                      out.data_ptr(), Cn, _stream(h)), "dgn_readout_forward")
        _count(1)
        ctx.graph, ctx.op = graph, op
        ctx.save_for_backward(h, out)
        return out

    @staticmethod
    def backward(ctx, g_out):
        h, out = ctx.saved_tensors
        g_out = g_out.contiguous()
        B, Cn = out.shape
        d_h = torch.empty_like(h)
        check(lib.dgn_readout_backward(B, ctx.graph.graph_ptr.data_ptr(), Cn, h.data_ptr(), h.stride(0),
                                       out.data_ptr(), Cn, ctx.op, g_out.data_ptr(), Cn, d_h.data_ptr(),
                                       d_h.stride(0), h.shape[0], _stream(h)), "dgn_readout_backward")
        _count(1)
        return d_h, None, None


def readout(graph, h, op: str):
    """Per-graph ``sum`` / ``mean`` / ``max`` over the node rows (dgl.{sum,mean,max}_nodes)."""
    code = {"sum": _lib.READOUT_SUM, "mean": _lib.READOUT_MEAN, "max": _lib.READOUT_MAX}[op]
    return _Readout.apply(h, graph, code)


_EMB_WS = {}


def _emb_workspace(vocab, cols, device):
    """Zero-initialised, reused workspace of dgn_embedding_backward (its counters return to zero)."""
    key = (vocab, cols, str(device))
    ws = _EMB_WS.get(key)
    if ws is None:
        ws = torch.zeros(32 * vocab * cols + 64, device=device, dtype=torch.float32)
        _EMB_WS[key] = ws
    return ws


class _Embedding(torch.autograd.Function):
    """nn.Embedding lookup whose weight gradient is the deterministic dgn_embedding_backward kernel."""

    @staticmethod
    def forward(ctx, weight, idx, n_rows_dev, direct):
        ctx.save_for_backward(idx)
        ctx.wref, ctx.n_rows_dev, ctx.direct = weight, n_rows_dev, direct
        return weight.index_select(0, idx)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        w = ctx.wref
        g = g.contiguous()
        direct = ctx.direct and w.grad is not None and w.grad.is_contiguous()
        dw = w.grad if direct else torch.zeros_like(w)
        nd = ctx.n_rows_dev
        ws = _emb_workspace(w.shape[0], g.shape[1], g.device)
        check(lib.dgn_embedding_backward(g.shape[0], g.shape[1], w.shape[0], idx.data_ptr(), g.data_ptr(),
                                         g.stride(0), dw.data_ptr(), dw.stride(0),
                                         nd.data_ptr() if nd is not None else None, ws.data_ptr(), _stream(g)),
              "dgn_embedding_backward")
        _count(1)
        return (None if direct else dw), None, None, None


def embedding(weight, idx, n_rows_dev=None, direct_grad=None):
    """``weight[idx]`` (rb/nets/molecules_graph_regression/dgn_net.py:58) with a deterministic backward.

    ``direct_grad=True`` accumulates straight into ``weight.grad`` (which must already exist) instead of
    returning a gradient tensor - the engine's flat gradient buffer makes autograd's extra add redundant.
    Falls back to torch's own embedding when the vocabulary does not fit the kernel's shared-memory tile."""
    _need_cuda(weight, idx)
    if direct_grad is None:
        direct_grad = DIRECT_GRADS
    # the kernel reads idx as a contiguous 1-D int64 array and weight as contiguous fp32
    if weight.shape[0] > 200 or idx.dim() != 1 or weight.dtype != torch.float32 or not weight.is_contiguous():
        return torch.nn.functional.embedding(idx, weight)      # library op (still on the GPU)
    if idx.dtype != torch.int64 or not idx.is_contiguous():
        idx = idx.long().contiguous()
    return _Embedding.apply(weight, idx, n_rows_dev, direct_grad)


def eig_flip_(eig, seed, step):
    """In-place entry-wise random sign flip of ``ndata['eig']`` on the device - the ``--flip`` augmentation of
    rb/train/train_molecules_graph_regression.py:29-33 (one Bernoulli(1/2) per ENTRY, like the reference)."""
    _need_cuda(eig)
    if eig.dtype != torch.float32 or not eig.is_contiguous():
        raise _lib.DgnError("eig_flip_ needs a contiguous float32 tensor")
    check(lib.dgn_eig_flip(eig.data_ptr(), eig.numel(), int(seed), int(step), _stream(eig)), "dgn_eig_flip")
    _count(1)
    try:                                   # tell torch (and BatchedGraph.field's staleness check) about the in-place change
        torch._C._autograd._unsafe_set_version_counter([eig], [eig._version + 1])
    except Exception:
        eig.add_(0)
    return eig


# ---------------------------------------------------------------------------------------------------------
# fp32-accurate tensor-core GEMM (tcgen05, 3xTF32)
# ---------------------------------------------------------------------------------------------------------
_GEMM_WS = {}
GEMM_ENABLED = True          # set False to route every GEMM through the library (A/B comparisons)


def _gemm_workspace(device):
    """Split-K workspace, one per (device, stream): GEMMs on the side stream run concurrently with the main one."""
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    ws = _GEMM_WS.get(key)
    if ws is None:
        ws = torch.zeros(int(lib.dgn_gemm_ws_floats()), device=device, dtype=torch.float32)
        _GEMM_WS[key] = ws
    return ws


def gemm(a, b, a_kmajor=True, b_kmajor=True, out=None, accumulate=False, c_transposed=False):
    """``C (+)= op(A) @ op(B)`` in fp32 accuracy on the tcgen05 tensor cores (``dgn_gemm_tf32x3``).

    ``a`` is ``[M, K]`` (``a_kmajor``) or ``[K, M]``; ``b`` is ``[N, K]`` (``b_kmajor``, i.e. ``A @ B.T``) or ``[K, N]``.
    ``out`` is ``[M, N]`` (or ``[N, M]`` with ``c_transposed``), row stride free, unit column stride.  Shapes the kernel
    does not take (rows not 16 B aligned) go through the library GEMM on the same device."""
    _need_cuda(a, b)
    M, K = (a.shape[0], a.shape[1]) if a_kmajor else (a.shape[1], a.shape[0])
    N = b.shape[0] if b_kmajor else b.shape[1]
    if out is None:
        out = torch.empty((N, M) if c_transposed else (M, N), device=a.device, dtype=torch.float32)
        accumulate = False
    # the tensor-core kernel wins once there is real work (long K or a large output); tiny products are
    # launch-latency bound and stay with the library (measured: tools/gemm_bench.py, profiles/README.md)
    worth = K >= 256 or M * N >= (1 << 20)
    ok = (GEMM_ENABLED and worth and a.dtype == torch.float32 and b.dtype == torch.float32 and a.stride(1) == 1 and
          b.stride(1) == 1 and out.stride(1) == 1 and K > 0)
    if ok:
        rc = lib.dgn_gemm_tf32x3(M, N, K, a.data_ptr(), a.stride(0), int(a_kmajor), b.data_ptr(), b.stride(0),
                                 int(b_kmajor), out.data_ptr(), out.stride(0), int(accumulate), int(c_transposed),
                                 _gemm_workspace(a.device).data_ptr(), _stream(a))
        if rc == 0:
            _count(1)
            return out
        if rc != -2:                      # anything but "unsupported shape / alignment" is an error
            check(rc, "dgn_gemm_tf32x3")
    A = a if a_kmajor else a.t()
    Bm = b.t() if b_kmajor else b
    if c_transposed:                              # (A @ B)^T = B^T @ A^T
        A, Bm = Bm.t(), A.t()
    if accumulate:
        out.addmm_(A, Bm)                         # one library launch, beta = 1
    else:
        torch.mm(A, Bm, out=out)
    return out


# ---------------------------------------------------------------------------------------------------------
# scaler-folded posttrans products (dgn_post_*): the [N, S*A*F] concatenation never exists
# ---------------------------------------------------------------------------------------------------------
FOLD_ENABLED = os.environ.get("DGN_NO_FOLD", "0") != "1"


class PostSpec:
    """Shapes / scalers of one layer's posttrans for ``dgn_post_*``: ``cat = [h (lead columns) | raw aggregates]``."""

    def __init__(self, spec: AggSpec, lead: int, n_out: int):
        self.lead, self.n_agg, self.n_out = int(lead), spec.A * spec.F, int(n_out)
        self.S_decl, self.S = spec.S_decl, spec.S
        self.kinds = [sc.kind for sc in spec.scalers]
        self.avg_log = spec.avg_log
        self.w_cols = self.lead + self.S * self.n_agg

    def supported(self, spec: AggSpec) -> bool:
        return (FOLD_ENABLED and spec.Fg == spec.F and self.lead % 4 == 0 and self.n_agg % 4 == 0 and
                self.n_out % 4 == 0 and self.w_cols % 4 == 0)

    def args(self, graph, cat, W):
        a = _lib.DgnPostArgs()
        a.n_rows, a.n_lead, a.n_agg, a.n_out, a.n_scalers = cat.shape[0], self.lead, self.n_agg, self.n_out, self.S_decl
        for i, k in enumerate(self.kinds):
            a.scaler_kind[i] = k
        a.avg_log = self.avg_log
        a.log_deg = graph.log_deg.data_ptr()
        a.cat, a.ld_cat, a.w, a.ld_w = cat.data_ptr(), cat.stride(0), W.data_ptr(), W.stride(0)
        return a


def post_forward(ps: PostSpec, graph, cat, W, y, stats=None, y_bias=None, snorm=None, n_rows_dev=None):
    """``y = h W_h^T + sum_s c_s (agg W_s^T)``.  With ``stats`` (the workspace of the norm call that follows) the kernel
    also leaves partial batch statistics of ``(y + y_bias) * snorm`` there, one slab per 128-row tile.  Returns
    ``stat_parts`` for ``norm_forward_raw``: 0 = nothing written, > 0 = that many slabs."""
    st, parts = None, C.c_int32(0)
    if stats is not None:
        st = _lib.DgnPostStats(stats.data_ptr(), y_bias.data_ptr() if y_bias is not None else None,
                               snorm.data_ptr() if snorm is not None else None,
                               n_rows_dev.data_ptr() if n_rows_dev is not None else None)
    check(lib.dgn_post_forward(C.byref(ps.args(graph, cat, W)), y.data_ptr(), y.stride(0),
                               C.byref(st) if st is not None else None, C.byref(parts), _stream(cat)),
          "dgn_post_forward")
    _count(1)
    return int(parts.value)


def post_backward(ps: PostSpec, graph, cat, W, d_y, d_cat):
    check(lib.dgn_post_backward(C.byref(ps.args(graph, cat, W)), d_y.data_ptr(), d_y.stride(0), d_cat.data_ptr(),
                                d_cat.stride(0), _stream(cat)), "dgn_post_backward")
    _count(1)


def post_wgrad(ps: PostSpec, graph, cat, W, d_y, d_w, accumulate):
    check(lib.dgn_post_wgrad(C.byref(ps.args(graph, cat, W)), d_y.data_ptr(), d_y.stride(0), d_w.data_ptr(),
                             d_w.stride(0), int(accumulate), _stream(cat)), "dgn_post_wgrad")
    _count(1)


def pre_wgrad(h, d_P, d_Q, d_w, d_b, accumulate):
    """``d_w[:, :Fi] (+)= d_P^T h ; d_w[:, Fi:2Fi] (+)= d_Q^T h ; d_b (+)= sum_rows d_Q`` in one launch; returns False
    when the shapes are outside the kernel's range (callers then use the generic GEMMs)."""
    rc = lib.dgn_pre_wgrad(h.shape[0], h.shape[1], d_P.shape[1], h.data_ptr(), h.stride(0), d_P.data_ptr(),
                           d_P.stride(0), d_Q.data_ptr(), d_Q.stride(0), d_w.data_ptr(), d_w.stride(0),
                           d_b.data_ptr() if d_b is not None else None, int(accumulate), _stream(h))
    if rc == -2:
        return False
    check(rc, "dgn_pre_wgrad")
    _count(1)
    return True


# ---------------------------------------------------------------------------------------------------------
# side stream for work that is off the critical path (weight-gradient GEMMs)
# ---------------------------------------------------------------------------------------------------------
class _SideQueue:
    """A second CUDA stream that forks from / joins the current one.  Under CUDA-graph capture the forked work
    becomes a parallel branch of the graph, so the small critical-path kernels and the weight-gradient GEMMs
    (only needed by the optimizer at the very end) share the GPU instead of running back to back."""

    def __init__(self, device):
        self.stream = torch.cuda.Stream(device=device)
        self.keep, self.dirty = [], False

    def run(self, fn, keep=()):
        main = torch.cuda.current_stream(self.stream.device)
        self.stream.wait_stream(main)                 # inputs produced so far on the main stream are ready
        with torch.cuda.stream(self.stream):
            fn()
        self.keep.extend(keep)                        # operands must outlive the forked work: hold them until join()
        self.dirty = True

    def join(self):
        if self.dirty:
            torch.cuda.current_stream(self.stream.device).wait_stream(self.stream)
            self.keep.clear()
            self.dirty = False


_SIDE = {}
SIDE_STREAM_ENABLED = False      # only inside step_scope(): plain autograd use stays single-stream
DIRECT_GRADS = False             # only inside step_scope(): parameter gradients accumulate straight into .grad
ACTIVE_SCOPE = None              # the step_scope in force (towers.py batches its packing copies per scope)


class step_scope:
    """What a step runner (engine.TrainStep) switches on for the duration of ITS OWN forward + backward, and nothing
    else: in-place accumulation of parameter gradients into existing ``.grad`` buffers (autograd receives None for
    them - that breaks ``torch.autograd.grad``, hooks and DDP, so it is opt-in), the forked stream for
    weight-gradient GEMMs, and the collected BatchNorm batch counters.  Leaving the scope joins the side stream and
    restores every switch, so a later ``backward()`` outside the runner is plain single-stream autograd again."""

    def __init__(self, device, direct_grads=True, side_stream=True):
        self.device, self.direct, self.side = device, direct_grads, side_stream
        self.tower_state = {}            # towers.py: which packed layers were synchronised / ran in this scope

    def __enter__(self):
        global SIDE_STREAM_ENABLED, DIRECT_GRADS, BN_COUNTERS, ACTIVE_SCOPE
        self._saved = (SIDE_STREAM_ENABLED, DIRECT_GRADS, BN_COUNTERS, ACTIVE_SCOPE)
        SIDE_STREAM_ENABLED, DIRECT_GRADS, BN_COUNTERS, ACTIVE_SCOPE = self.side, self.direct, [], self
        return self

    def flush_bn_counters(self):
        global BN_COUNTERS
        if BN_COUNTERS:
            torch._foreach_add_(BN_COUNTERS, 1)
        BN_COUNTERS = []

    def __exit__(self, *exc):
        global SIDE_STREAM_ENABLED, DIRECT_GRADS, BN_COUNTERS, ACTIVE_SCOPE
        side_join(self.device)
        if self.tower_state and exc[0] is None:
            from . import towers
            towers.finish_scope(self)    # packed layers: running statistics and gradients back to the modules (2 launches)
        SIDE_STREAM_ENABLED, DIRECT_GRADS, BN_COUNTERS, ACTIVE_SCOPE = self._saved
        return False


def side_queue(device):
    if not SIDE_STREAM_ENABLED:
        return None
    q = _SIDE.get(str(device))
    if q is None:
        q = _SideQueue(device)
        _SIDE[str(device)] = q
    return q


def side_join(device):
    q = _SIDE.get(str(device))
    if q is not None:
        q.join()


def pair_linear_forward(h, W, Fi):
    """``P = h @ W[:, :Fi].T``, ``Q = h @ W[:, Fi:2Fi].T`` in one launch (dgn_pair_linear_forward); library GEMMs
    when the shape is outside the kernel's range."""
    N, Fo = h.shape[0], W.shape[0]
    P = torch.empty((N, Fo), device=h.device, dtype=torch.float32)
    Q = torch.empty((N, Fo), device=h.device, dtype=torch.float32)
    rc = -2
    if N > 0 and h.stride(1) == 1 and W.stride(1) == 1:
        rc = lib.dgn_pair_linear_forward(N, Fi, Fo, h.data_ptr(), h.stride(0), W.data_ptr(), W.stride(0), P.data_ptr(),
                                         P.stride(0), Q.data_ptr(), Q.stride(0), _stream(h))
    if rc == -2:
        torch.mm(h, W[:, :Fi].t(), out=P)
        torch.mm(h, W[:, Fi:2 * Fi].t(), out=Q)
        return P, Q
    check(rc, "dgn_pair_linear_forward")
    _count(1)
    return P, Q


def pair_gather_supported(Fi, Fo, *tensors):
    return (FOLD_ENABLED and 0 < Fi <= 128 and 0 < Fo <= 128 and Fi % 4 == 0 and Fo % 4 == 0 and
            all(t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0 for t in tensors))


def pair_gather_backward(graph, edge_ws, d_Q, W, Fi, d_h, d_P):
    """``d_P[u] = sum_{out-edges} edge_ws[slot]`` ; ``d_h += d_P W[:, :Fi] + d_Q W[:, Fi:2Fi]`` ; ``d_P`` written - the
    source-side reduction of the aggregation backward and the pretrans input gradient in one launch."""
    N, Fo = d_Q.shape
    check(lib.dgn_pair_gather_backward(N, Fi, Fo, graph.out_ptr.data_ptr(), graph.out_slot.data_ptr(),
                                       edge_ws.data_ptr(), edge_ws.stride(0), d_Q.data_ptr(), d_Q.stride(0),
                                       W.data_ptr(), W.stride(0), d_h.data_ptr(), d_h.stride(0), d_P.data_ptr(),
                                       d_P.stride(0), _stream(d_h)), "dgn_pair_gather_backward")
    _count(1)


def pair_linear_backward(d_P, d_Q, W, Fi, d_h):
    """``d_h += d_P @ W[:, :Fi] + d_Q @ W[:, Fi:2Fi]`` in one launch (dgn_pair_linear_backward)."""
    N, Fo = d_P.shape
    rc = -2
    if N > 0 and W.stride(1) == 1:
        rc = lib.dgn_pair_linear_backward(N, Fi, Fo, d_P.data_ptr(), d_P.stride(0), d_Q.data_ptr(), d_Q.stride(0),
                                          W.data_ptr(), W.stride(0), d_h.data_ptr(), d_h.stride(0), _stream(d_h))
    if rc == -2:
        d_h.addmm_(d_P, W[:, :Fi])
        d_h.addmm_(d_Q, W[:, Fi:2 * Fi])
        return d_h
    check(rc, "dgn_pair_linear_backward")
    _count(1)
    return d_h


# ---------------------------------------------------------------------------------------------------------
# graph-level prediction head (MLPReadout, L = 2) and L1 loss: one launch per direction each
# ---------------------------------------------------------------------------------------------------------
HEAD_ENABLED = os.environ.get("DGN_NO_HEAD", "0") != "1"


def _head_args(x, w1, b1, w2, b2, w3, b3, a1, a2, y):
    a = _lib.DgnHeadArgs()
    a.n_rows, a.d0, a.d1, a.d2, a.d_out = x.shape[0], w1.shape[1], w1.shape[0], w2.shape[0], w3.shape[0]
    a.x, a.ld_x = x.data_ptr(), x.stride(0)
    a.w1, a.b1, a.w2, a.b2, a.w3, a.b3 = (t.data_ptr() for t in (w1, b1, w2, b2, w3, b3))
    a.a1, a.a2, a.y, a.ld_y = a1.data_ptr(), a2.data_ptr(), y.data_ptr(), y.stride(0)
    return a


def head_supported(x, fcs) -> bool:
    """Shapes dgn_head_forward takes: 3 Linear layers with bias on CUDA fp32, sizes within the kernel's limits."""
    if not (HEAD_ENABLED and x.is_cuda and x.dim() == 2 and len(fcs) == 3 and x.dtype == torch.float32):
        return False
    if any(fc.bias is None or not fc.weight.is_contiguous() or fc.weight.dtype != torch.float32 for fc in fcs):
        return False
    d0, d1, d2, do = fcs[0].in_features, fcs[0].out_features, fcs[1].out_features, fcs[2].out_features
    return (x.shape[0] <= 1024 and x.shape[1] == d0 and d1 * d0 <= 4096 and d2 * d1 <= 1024 and do * d2 <= 1024 and
            d1 + d2 + do <= 1024 and (d1 * d0 + d2 * d1 + do * d2 + 32 * (d0 + 2 * d1 + 2 * d2 + do)) * 4 <= 150 * 1024)


class _Head(torch.autograd.Function):
    """y = W3 relu(W2 relu(W1 x + b1) + b2) + b3 (rb/nets/mlp_readout_layer.py:24-30) through dgn_head_forward /
    dgn_head_backward.  With ``direct`` and existing ``.grad`` buffers the parameter gradients are accumulated in
    place (no autograd accumulation kernels), like the fused layer."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, w3, b3, direct):
        x = _f32c(x)
        B = x.shape[0]
        a1 = torch.empty((B, w1.shape[0]), device=x.device, dtype=torch.float32)
        a2 = torch.empty((B, w2.shape[0]), device=x.device, dtype=torch.float32)
        y = torch.empty((B, w3.shape[0]), device=x.device, dtype=torch.float32)
        if B > 0:
            check(lib.dgn_head_forward(C.byref(_head_args(x, w1, b1, w2, b2, w3, b3, a1, a2, y)), _stream(x)),
                  "dgn_head_forward")
            _count(1)
        ctx.save_for_backward(x, w1, b1, w2, b2, w3, b3, a1, a2, y)
        ctx.direct = direct
        return y

    @staticmethod
    def backward(ctx, g_y):
        x, w1, b1, w2, b2, w3, b3, a1, a2, y = ctx.saved_tensors
        params = (w1, b1, w2, b2, w3, b3)
        g_y = g_y.contiguous()
        direct = ctx.direct and all(p.grad is not None and p.grad.is_contiguous() for p in params)
        grads = [p.grad for p in params] if direct else [torch.empty_like(p) for p in params]
        d_x = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        if x.shape[0] > 0:
            g = _lib.DgnHeadGrad()
            g.g_y, g.ld_gy = g_y.data_ptr(), g_y.stride(0)
            if d_x is not None:
                g.d_x, g.ld_dx = d_x.data_ptr(), d_x.stride(0)
            g.d_w1, g.d_b1, g.d_w2, g.d_b2, g.d_w3, g.d_b3 = (t.data_ptr() for t in grads)
            g.accumulate = int(direct)
            check(lib.dgn_head_backward(C.byref(_head_args(x, w1, b1, w2, b2, w3, b3, a1, a2, y)), C.byref(g),
                                        _stream(x)), "dgn_head_backward")
            _count(1)
        elif not direct:
            grads = [torch.zeros_like(p) for p in params]
        if direct:
            return (d_x,) + (None,) * 7
        return (d_x,) + tuple(grads) + (None,)


def mlp_head(x, fcs, direct_grads=None):
    """MLPReadout forward over its three ``nn.Linear`` layers in one launch (one more for the backward).
    ``direct_grads`` defaults to the enclosing ``step_scope`` (off for plain autograd use)."""
    if direct_grads is None:
        direct_grads = DIRECT_GRADS
    return _Head.apply(x, fcs[0].weight, fcs[0].bias, fcs[1].weight, fcs[1].bias, fcs[2].weight, fcs[2].bias,
                       direct_grads)


class _L1Loss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, target):
        y, target = y.contiguous(), target.contiguous()
        loss = torch.empty((), device=y.device, dtype=torch.float32)
        check(lib.dgn_l1_loss_forward(y.numel(), y.data_ptr(), target.data_ptr(), loss.data_ptr(), _stream(y)),
              "dgn_l1_loss_forward")
        _count(1)
        ctx.save_for_backward(y, target)
        return loss

    @staticmethod
    def backward(ctx, g):
        y, target = ctx.saved_tensors
        g = g.contiguous().float()
        d_y = torch.empty_like(y)
        check(lib.dgn_l1_loss_backward(y.numel(), y.data_ptr(), target.data_ptr(), g.data_ptr(), d_y.data_ptr(),
                                       _stream(y)), "dgn_l1_loss_backward")
        _count(1)
        return d_y, None


def l1_loss(scores, targets):
    """``nn.L1Loss()(scores, targets)`` (rb/nets/molecules_graph_regression/dgn_net.py:90-92): one launch per direction
    on CUDA fp32 tensors of equal shape; anything else goes through torch."""
    if (HEAD_ENABLED and scores.is_cuda and scores.dtype == torch.float32 and targets.dtype == torch.float32 and
            scores.shape == targets.shape and scores.numel() > 0 and not targets.requires_grad):
        return _L1Loss.apply(scores, targets)
    return torch.nn.functional.l1_loss(scores, targets)
