"""Whole-layer autograd node for the default DGN layer configuration.

``pretrans_layers == posttrans_layers == 1`` (all five reference configs, rb/configs/*.json) makes a DGN
layer a fixed chain

    P = h W_src^T, Q = h W_dst^T                         1 launch (node level), dgn_pair_linear_forward
    cat = [h | aggregators(P[u] + Q[v] + b)]             dgn_agg_forward      (b fused as q_bias; RAW aggregates)
    y = h W_h^T + sum_s c_s(v) (agg W_s^T)               dgn_post_forward     (scalers folded into the GEMM epilogue:
                                                                               the [N, S*A*F] concatenation of
                                                                               rb/nets/dgn_layer.py:94-96 never exists)
    out = relu(BN((y + b_post) * snorm_n)) + h           dgn_norm_forward     (b_post fused as y_bias)

(shapes the folded kernels do not take - widths that are not multiples of 4, towers in one launch - fall back to the
scaled ``[N, (1 + S*A) F]`` concatenation and the generic ``dgn_gemm_tf32x3``.)

Running that chain through generic autograd costs ~50 kernels per layer and direction in glue: slice
backward, gradient accumulation adds, bias-gradient reductions, fills.  This node owns the whole chain
instead: 6 launches forward, ~11 backward, every GEMM writes (beta = 1) straight into its destination and -
when the parameters already own ``.grad`` buffers, as under ``engine.TrainStep`` - parameter gradients are
accumulated in place so autograd has nothing left to add.

Same arithmetic as the op-level path in ``nets/dgn_layer.py`` (tests compare both against the golden
vectors of the reference).
"""
from __future__ import annotations

import torch

from . import _lib, ops
from .ops import (agg_backward_raw, agg_forward_raw, gemm, norm_backward_raw, norm_forward_raw, norm_pair_forward,
                  pair_linear_backward, pair_linear_forward, post_backward, post_forward, post_wgrad, pre_wgrad, side_queue, _f32c, _need_cuda)

_ONES = {}


def _ones(n, device):
    key = (n, str(device))
    t = _ONES.get(key)
    if t is None:
        t = torch.ones(n, device=device, dtype=torch.float32)
        _ONES[key] = t
    return t


class LayerConfig:
    """Non-tensor state of one fused layer call."""
    __slots__ = ("graph", "spec", "eig", "snorm", "bn", "training", "relu", "residual", "direct", "in_dim",
                 "has_pretrans", "params", "aspec", "post", "pq", "next_w", "pq_out")

    def __init__(self, graph, spec, eig, snorm, bn, training, relu, residual, direct, in_dim, has_pretrans, params,
                 spec_raw=None, post=None, pq=None, next_w=None):
        # pq: (P, Q) of THIS layer already computed by the previous layer's epilogue; next_w: pretrans weight of the
        # NEXT layer, whose P / Q this layer's epilogue computes (returned through pq_out)
        self.pq, self.next_w, self.pq_out = pq, next_w, None
        self.graph, self.spec, self.eig, self.snorm, self.bn = graph, spec, eig, snorm, bn
        self.training, self.relu, self.residual, self.direct = training, relu, residual, direct
        self.in_dim, self.has_pretrans, self.params = in_dim, has_pretrans, params
        # folded path: the aggregation runs with the single-scaler spec (raw aggregates), `post` folds the scalers
        self.post = post if (post is not None and spec_raw is not None) else None
        self.aspec = spec_raw if self.post is not None else spec


class _FusedLayer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, h, R, W_pre, b_pre, W_post, b_post, gamma, beta):
        _need_cuda(h)
        h = _f32c(h)
        g, spec, Fi = cfg.graph, cfg.aspec, cfg.in_dim
        N, dev = h.shape[0], h.device
        P = Q = None
        if cfg.has_pretrans:
            if cfg.pq is not None:
                P, Q = cfg.pq                            # computed by the previous layer's fused epilogue
            else:
                P, Q = pair_linear_forward(h, W_pre, Fi)     # h @ W_src^T, h @ W_dst^T in one launch
            cat = torch.empty((N, Fi + spec.out_width), device=dev, dtype=torch.float32)
            agg_forward_raw(g, spec, _lib.MSG_AFFINE, P, Q, R, h, cfg.eig, cat, True, q_bias=b_pre)
        else:                                            # simple layer: message = h[src], no h block
            cat = torch.empty((N, spec.out_width), device=dev, dtype=torch.float32)
            agg_forward_raw(g, spec, _lib.MSG_SOURCE, h, None, None, h, cfg.eig, cat, False)
        Co = W_post.shape[0]
        out = torch.empty((N, Co), device=dev, dtype=torch.float32)
        stats = torch.empty(_lib.NORM_WS_PER_COL * Co, device=dev, dtype=torch.float32)
        bn = cfg.bn
        use_batch = True
        if bn is not None:
            if cfg.training:
                ops.count_bn_batch(bn)
            use_batch = cfg.training or not bn.track_running_stats
        stat_parts = 0
        if cfg.post is not None:                         # h W_h^T + sum_s c_s (agg W_s^T): scalers folded in the epilogue
            y = torch.empty((N, Co), device=dev, dtype=torch.float32)
            if N > 0:
                want_stats = bn is not None and use_batch      # BatchNorm partial statistics as a by-product
                stat_parts = post_forward(cfg.post, g, cat, W_post, y, stats if want_stats else None, b_post, cfg.snorm,
                                          getattr(g, "n_rows_dev", None))
        else:
            y = gemm(cat, W_post)                        # cat @ W_post^T
        # epilogue, fused with the next layer's node-level pretrans halves when that layer is known and the batch
        # statistics are final (or not needed)
        fuse_next = (cfg.next_w is not None and N > 0 and cfg.next_w.shape[1] >= 2 * Co and
                     (bn is None or not use_batch or stat_parts > 0))
        nargs = norm_forward_raw(
            y, out, stats, snorm=cfg.snorm, y_bias=b_post, launch=not fuse_next,
            gamma=gamma if bn is not None else None, beta=beta if bn is not None else None,
            running_mean=bn.running_mean if bn is not None else None,
            running_var=bn.running_var if bn is not None else None,
            momentum=(0.1 if bn is None or bn.momentum is None else bn.momentum),
            eps=(1e-5 if bn is None else bn.eps), training=use_batch, relu=cfg.relu,
            residual=h if cfg.residual else None, n_rows_dev=getattr(g, "n_rows_dev", None), stat_parts=stat_parts)
        if fuse_next:
            Fn = cfg.next_w.shape[0]
            Pn = torch.empty((N, Fn), device=dev, dtype=torch.float32)
            Qn = torch.empty((N, Fn), device=dev, dtype=torch.float32)
            if norm_pair_forward(nargs, cfg.next_w, Pn, Qn):
                cfg.pq_out = (Pn, Qn)
            else:                                        # shapes outside the fused kernel: plain epilogue
                ops.check(ops.lib.dgn_norm_forward(ops.C.byref(nargs), ops._stream(y)), "dgn_norm_forward")
                ops._count(1)
        ctx.cfg, ctx.nargs = cfg, nargs
        ctx.save_for_backward(h, R, P, Q, cat, y, stats, W_pre, b_pre, W_post, b_post, gamma, beta)
        return out

    @staticmethod
    def backward(ctx, g_out):
        cfg = ctx.cfg
        h, R, P, Q, cat, y, stats, W_pre, b_pre, W_post, b_post, gamma, beta = ctx.saved_tensors
        g, spec, Fi = cfg.graph, cfg.aspec, cfg.in_dim
        N, E, dev = h.shape[0], g.number_of_edges(), h.device
        g_out = g_out.contiguous()
        Co = y.shape[1]
        pW_pre, pb_pre, pW_post, pb_post, pgamma, pbeta = cfg.params
        direct = cfg.direct and all(p is None or p.grad is not None for p in cfg.params)

        # ---- norm / BN / ReLU / residual ---------------------------------------------------------------------------
        d_y = torch.empty_like(y)
        scratch = torch.empty(_lib.NORM_WS_PER_COL * Co, device=dev, dtype=torch.float32)
        has_bn = cfg.bn is not None
        if direct:
            d_gamma, d_beta, d_bpost = (pgamma.grad if has_bn else None), (pbeta.grad if has_bn else None), pb_post.grad
        else:
            d_gamma = torch.empty(Co, device=dev) if has_bn else None
            d_beta = torch.empty(Co, device=dev) if has_bn else None
            d_bpost = torch.empty(Co, device=dev)
        norm_backward_raw(ctx.nargs, g_out, d_y, scratch, d_gamma, d_beta, d_bpost, accumulate=direct)
        fold = cfg.post is not None and N > 0
        if fold:
            d_cat = torch.empty_like(cat)                # [d_y W_h | sum_s c_s (d_y W_s)]: gradient of [h | raw aggregates]
            post_backward(cfg.post, g, cat, W_post, d_y, d_cat)
        else:
            d_cat = gemm(d_y, W_post, b_kmajor=False)    # d_y @ W_post
        # dW_post = d_y^T @ cat, computed as (cat^T @ d_y)^T so the 128-row tile dimension is the wide one.
        # Weight gradients are off the critical path: with direct accumulation they run on the side stream.
        side = side_queue(dev) if direct else None
        if direct:
            def _dw_post():
                if fold:
                    post_wgrad(cfg.post, g, cat, W_post, d_y, pW_post.grad, True)
                else:
                    gemm(cat, d_y, a_kmajor=False, b_kmajor=False, out=pW_post.grad, accumulate=True, c_transposed=True)
            if side is not None:
                side.run(_dw_post, keep=(cat, d_y))
            else:
                _dw_post()
            d_Wpost = None
        elif fold:
            d_Wpost = torch.empty_like(W_post)
            post_wgrad(cfg.post, g, cat, W_post, d_y, d_Wpost, False)
        else:
            d_Wpost = gemm(cat, d_y, a_kmajor=False, b_kmajor=False, c_transposed=True)

        # ---- aggregation -----------------------------------------------------------------------------------------------
        d_h = torch.empty((N, Fi), device=dev, dtype=torch.float32)
        ws = torch.empty((max(E, 1), Fi), device=dev, dtype=torch.float32)
        d_R = d_Wpre = d_bpre = None
        resid = g_out if cfg.residual else None
        if cfg.has_pretrans:
            d_P = torch.empty((N, Fi), device=dev, dtype=torch.float32)
            d_Q = torch.empty((N, Fi), device=dev, dtype=torch.float32)
            if R is not None:
                # zeros: only rows of real edges are written; the padding rows of a fixed-capacity batch flow into
                # dW_e = d_R^T @ e (and the pretrans MLP backward) and must not carry allocator garbage
                d_R = torch.zeros((max(E, 1), Fi), device=dev, dtype=torch.float32)[:E]
            if N > 0 and W_pre.stride(1) == 1 and ops.pair_gather_supported(Fi, d_P.shape[1], ws, d_Q, W_pre, d_h, d_P):
                # the per-edge message gradients stay in `ws`; their source-side reduction happens inside the
                # pretrans input-gradient kernel (one launch instead of two)
                agg_backward_raw(g, spec, _lib.MSG_AFFINE, P, Q, R, h, cfg.eig, d_cat, True, d_x=None, d_q=d_Q, d_r=d_R,
                                 d_h=d_h, edge_ws=ws, q_bias=b_pre, d_h_addend=resid)
                ops.pair_gather_backward(g, ws, d_Q, W_pre, Fi, d_h, d_P)
            else:
                agg_backward_raw(g, spec, _lib.MSG_AFFINE, P, Q, R, h, cfg.eig, d_cat, True, d_x=d_P, d_q=d_Q, d_r=d_R,
                                 d_h=d_h, edge_ws=ws, q_bias=b_pre, d_h_addend=resid)
                pair_linear_backward(d_P, d_Q, W_pre, Fi, d_h)                               # += d_P W_src + d_Q W_dst
            if direct:
                gW = pW_pre.grad
                ones = _ones(N, dev)

                def _dw_pre():
                    if ops.FOLD_ENABLED and pre_wgrad(h, d_P, d_Q, gW, pb_pre.grad, True):   # one launch: both halves + bias
                        return
                    gemm(d_P, h, a_kmajor=False, b_kmajor=False, out=gW[:, :Fi], accumulate=True)          # += d_P^T @ h
                    gemm(d_Q, h, a_kmajor=False, b_kmajor=False, out=gW[:, Fi:2 * Fi], accumulate=True)    # += d_Q^T @ h
                    pb_pre.grad.addmv_(d_Q.t(), ones)
                # with edge features autograd itself accumulates into W_pre.grad on the main stream (backward of
                # R = e @ W_pre[:, 2F:]^T, a read-modify-write of the whole tensor): keep these on the main stream too
                if side is not None and R is None:
                    side.run(_dw_pre, keep=(d_P, d_Q, h))
                else:
                    _dw_pre()
            else:
                d_Wpre = torch.zeros_like(W_pre)          # columns past 2F (edge features) get theirs via R
                d_bpre = torch.empty_like(b_pre)
                if not (ops.FOLD_ENABLED and N > 0 and pre_wgrad(h, d_P, d_Q, d_Wpre, d_bpre, False)):
                    gemm(d_P, h, a_kmajor=False, b_kmajor=False, out=d_Wpre[:, :Fi])
                    gemm(d_Q, h, a_kmajor=False, b_kmajor=False, out=d_Wpre[:, Fi:2 * Fi])
                    d_bpre = torch.mv(d_Q.t(), _ones(N, dev))
        else:
            # x and h_in are the same tensor: the kernel folds d_h_in (+ residual) into the scattered gradient
            d_x = torch.empty((N, Fi), device=dev, dtype=torch.float32)
            agg_backward_raw(g, spec, _lib.MSG_SOURCE, h, None, None, h, cfg.eig, d_cat, False, d_x=d_x, d_h=d_h,
                             edge_ws=ws, fold_h_in=True, d_h_addend=resid)
            d_h = d_x
        if direct:
            return None, d_h, d_R, None, None, None, None, None, None
        return None, d_h, d_R, d_Wpre, d_bpre, d_Wpost, d_bpost, d_gamma, d_beta


def fused_layer(graph, spec, eig, h, R, pretrans_lin, posttrans_lin, bn, snorm, training, relu, residual, in_dim,
                direct_grads=None, spec_raw=None, post=None, pq=None, next_w=None):
    """One DGN layer (complex / tower when ``pretrans_lin`` is given, simple otherwise) as a single autograd node.

    ``direct_grads``: accumulate parameter gradients in place when every parameter already has ``.grad``; defaults
    to the enclosing ``ops.step_scope`` (off for plain autograd use, where autograd receives ordinary gradients)."""
    if direct_grads is None:
        direct_grads = ops.DIRECT_GRADS
    has_pre = pretrans_lin is not None
    W_pre = pretrans_lin.weight if has_pre else None
    b_pre = pretrans_lin.bias if has_pre else None
    gamma = bn.weight if bn is not None else None
    beta = bn.bias if bn is not None else None
    params = (W_pre, b_pre, posttrans_lin.weight, posttrans_lin.bias, gamma, beta)
    if snorm is not None:
        snorm = snorm.reshape(-1)
        if snorm.dtype != torch.float32:
            snorm = snorm.float()
        if not snorm.is_contiguous():
            snorm = snorm.contiguous()
    eig = _f32c(eig)
    cfg = LayerConfig(graph, spec, eig, snorm, bn, training, relu, residual, direct_grads, in_dim, has_pre, params,
                      spec_raw, post, pq, next_w)
    out = _FusedLayer.apply(cfg, h, R, W_pre, b_pre, posttrans_lin.weight, posttrans_lin.bias, gamma, beta)
    return out, cfg.pq_out
