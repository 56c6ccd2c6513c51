"""Single-launch tower layer: all T towers of a ``DGNLayerTower`` in ONE fused layer call - and, with T = 1, the padded
fast path for layer widths that are not multiples of 4 floats (the reference's own configs use 45 / 47 / 65 / 70:
rb/configs/*.json): the layer OWNS zero-padded operands (45 -> 48 columns) so that every row is 16 B aligned, the
128-bit row kernels and the tcgen05 GEMMs are taken, and the padding columns stay exactly zero through the layer.

The reference runs its towers one after the other (rb/nets/dgn_layer.py:309-325): T x (pretrans -> update_all -> posttrans ->
graph norm -> BatchNorm).  Every aggregator is column-wise independent and BatchNorm is per column, so the T towers over
feature slices of width F_t are exactly ONE layer of width F = T * F_t whose weights are block structured:

    W_pre  [F, 2F]          rows / columns of tower t only  (block diagonal, [src | dst] halves)
    W_post [F_o, (1+S*A) F] row block t reads column slice t of h and of every aggregate block

``TowerFusion`` packs the tower parameters into such dense operands with one ``dgn_segment_copy`` launch (a device table
of rectangular segments built once), runs the ordinary fused layer (``fused._FusedLayer``: one aggregation launch, one
posttrans GEMM, one epilogue for all towers) and scatters the weight gradients / BatchNorm running statistics back with
one launch each.  The zero blocks cost ~T x redundant flops in GEMMs that are latency bound at these sizes.
"""
from __future__ import annotations

import types
import weakref

import numpy as np
import torch

from . import _lib, ops
from .fused import LayerConfig, _FusedLayer

_SEG = np.dtype([("a", "<u8"), ("b", "<u8"), ("rows", "<i4"), ("cols", "<i4"), ("ld_a", "<i4"), ("ld_b", "<i4")])


class _Ctx:
    """Minimal stand-in for an autograd context, so that ``_FusedLayer.forward / backward`` can be driven directly."""

    def save_for_backward(self, *tensors):
        self.saved_tensors = tensors


def _launch(table, direction, device):
    _lib.check(_lib.lib.dgn_segment_copy(table.data_ptr(), table.numel() // _SEG.itemsize, direction,
                                         torch.cuda.current_stream(device).cuda_stream), "dgn_segment_copy")
    ops._count(1)


# ---- per-step batching of the packing copies -------------------------------------------------------------------------
# Inside a step runner's scope (ops.step_scope, i.e. engine.TrainStep) the packing copies of ALL packed layers of the
# device run as 4 launches per step instead of 5 per layer: [parameters + running statistics -> dense operands] and one
# multi-tensor zero of the dense gradients when the first packed layer of the step runs, [running statistics back] and
# [dense gradients += into .grad] when the scope closes.  The merged segment tables are cached per set of layers.
_LIVE = weakref.WeakSet()
_MERGED = {}
_SERIAL = [0]


def _merged(kind, fusions):
    key = (kind,) + tuple((id(f), f._serial) for f in fusions)
    t = _MERGED.get(key)
    if t is None:
        if len(_MERGED) > 64:
            _MERGED.clear()
        parts = []
        for f in fusions:
            for k in kind.split("+"):
                tab = getattr(f, k)
                if tab is not None and tab.numel():
                    parts.append(tab)
        t = torch.cat(parts) if parts else False
        _MERGED[key] = t
    return t


def _sync_in(scope, dev):
    """First packed layer of the step: refresh the dense operands of every prepared packed layer on `dev`."""
    st = scope.tower_state
    live = sorted((f for f in _LIVE if f._key is not None and f._key[0] == str(dev)), key=lambda f: f._order)
    for f in live:
        f.prepare(dev)                                            # re-validates the pointers the tables were built for
    live = [f for f in live if f.t_grads_direct is not None]
    tab = _merged("t_params+t_running", live)
    if tab is not False:
        _launch(tab, 0, dev)
    if live:
        torch._foreach_zero_([f.dense_grad for f in live])
    st["synced"] = {id(f) for f in live}
    st["fwd"], st["bwd"] = [], []


def finish_scope(scope):
    st, dev = scope.tower_state, scope.device
    fwd = [f for f in st.get("fwd", []) if f.bn is not None]
    if fwd:
        tab = _merged("t_running", fwd)
        if tab is not False:
            _launch(tab, 1, dev)                                  # updated running statistics back to the towers
    if st.get("bwd"):
        tab = _merged("t_grads_direct", st["bwd"])
        if tab is not False:
            _launch(tab, 2, dev)                                  # dense gradients += into the towers' .grad
    scope.tower_state = {}


class _TowerFused(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fusion, g, eig, snorm, training, h, *tower_params):
        f = fusion
        dev = h.device
        f.prepare(dev)
        scope = ops.ACTIVE_SCOPE if (ops.ACTIVE_SCOPE is not None and ops.DIRECT_GRADS and training) else None
        if scope is not None and "synced" not in scope.tower_state:
            _sync_in(scope, dev)
        batched = scope is not None and id(f) in scope.tower_state["synced"]
        has_bn = f.bn is not None
        if batched:
            scope.tower_state["fwd"].append(f)
        else:
            _launch(f.t_params, 0, dev)                           # tower parameters -> dense block operands (1 launch)
            if has_bn:
                _launch(f.t_running, 0, dev)                      # running statistics -> concatenated buffers
        direct = ops.DIRECT_GRADS and all(p.grad is not None for p in tower_params)
        cfg = LayerConfig(g, f.spec(eig.shape[1]), eig, snorm, f.bn, training, f.relu, f.residual, True, f.F, True,
                          (f.W_pre, f.b_pre, f.W_post, f.b_post, f.gamma if has_bn else None, f.beta if has_bn else None),
                          *f.folded(eig.shape[1]))
        inner = _Ctx()
        out = _FusedLayer.forward(inner, cfg, h, None, f.W_pre, f.b_pre, f.W_post, f.b_post,
                                  f.gamma if has_bn else None, f.beta if has_bn else None)
        if has_bn and training:
            if not batched:
                _launch(f.t_running, 1, dev)                      # updated running statistics back to the towers
            for bn in f.tower_bns:
                ops.count_bn_batch(bn)
        ctx.fusion, ctx.inner, ctx.direct, ctx.n_params = f, inner, direct, len(tower_params)
        ctx.scope = scope if (batched and direct) else None
        return out

    @staticmethod
    def backward(ctx, g_out):
        f, dev = ctx.fusion, g_out.device
        batched = ctx.scope is not None and ctx.scope is ops.ACTIVE_SCOPE and id(f) in ctx.scope.tower_state.get("synced", ())
        if not batched:
            f.dense_grad.zero_()                                  # the fused layer accumulates into these (direct mode)
        res = _FusedLayer.backward(ctx.inner, g_out)
        d_h = res[1]
        if batched:                                               # scattered with the other layers when the scope closes
            ctx.scope.tower_state["bwd"].append(f)
            return (None,) * 5 + (d_h,) + (None,) * ctx.n_params
        side = ops.side_queue(dev)
        if ctx.direct:
            def _scatter():
                _launch(f.t_grads_direct, 2, dev)                 # dense gradients += into the towers' .grad (1 launch)
            if side is not None:
                side.run(_scatter)                                # after the weight-gradient GEMMs on the same queue
            else:
                _scatter()
            return (None,) * 5 + (d_h,) + (None,) * ctx.n_params
        if side is not None:
            side.join()
        _launch(f.t_grads_stage, 1, dev)
        return (None,) * 5 + (d_h,) + tuple(t.clone() for t in f.stage_views)


class TowerFusion:
    """Packing state of one ``DGNLayerTower`` (built lazily on the first fused call)."""

    def __init__(self, convs, in_width, out_width, relu=False, residual=False):
        """``convs``: the T conv modules (``DGNTower``s, or one ``DGNLayerComplex``) of per-tower widths
        ``in_width -> out_width``; ``relu`` / ``residual``: the epilogue of the fused layer (towers have neither)."""
        self.convs = list(convs)
        self.Ft, self.Fo_t = int(in_width), int(out_width)
        self.Ftp, self.Fop = (self.Ft + 3) // 4 * 4, (self.Fo_t + 3) // 4 * 4        # layer-owned padded widths
        self.relu, self.residual = bool(relu), bool(residual)
        self._key = None
        self._specs = {}
        self._serial = 0                                           # bumped whenever prepare() rebuilds the tables
        _SERIAL[0] += 1
        self._order = _SERIAL[0]

    # ---- eligibility -------------------------------------------------------------------------------------------
    def supported(self, h) -> bool:
        t0 = self.convs[0]
        if not (ops.FOLD_ENABLED and h.is_cuda and not getattr(t0, "edge_features", False)):
            return False
        if t0.dropout and t0.training:
            return False
        for tw in self.convs:
            if not hasattr(tw, "pretrans") or tw._affine(tw.pretrans) is None or tw._affine(tw.posttrans) is None:
                return False
        if self.residual and self.Ftp != self.Fop:
            return False
        return self.Ftp * len(self.convs) <= 512 and self.Fop * len(self.convs) <= 512

    # ---- specs ---------------------------------------------------------------------------------------------------
    def spec(self, n_eig):
        sp = self._specs.get(n_eig)
        if sp is None:
            t0 = self.convs[0]
            sp = ops.AggSpec(t0.aggregators, t0.scalers, float(t0._spec(n_eig).avg_log), self.F, n_eig)
            self._specs[n_eig] = sp
        return sp

    def folded(self, n_eig):
        key = ("fold", n_eig)
        ent = self._specs.get(key)
        if ent is None:
            from .nets.scalers import SCALERS
            spec = self.spec(n_eig)
            post = ops.PostSpec(spec, self.F, self.Fo)
            if post.supported(spec):
                ent = (ops.AggSpec(spec.aggregators, [SCALERS["identity"]], spec.avg_log, self.F, n_eig), post)
            else:
                ent = (None, None)
            self._specs[key] = ent
        return ent

    # ---- dense operands and segment tables ----------------------------------------------------------------------
    def _tower_tensors(self):
        pre = [tw._affine(tw.pretrans) for tw in self.convs]
        post = [tw._affine(tw.posttrans) for tw in self.convs]
        bns = [tw.batchnorm_h for tw in self.convs] if self.convs[0].batch_norm else None
        params = []
        for t in range(len(self.convs)):
            params += [pre[t].weight, pre[t].bias, post[t].weight, post[t].bias]
            if bns is not None:
                params += [bns[t].weight, bns[t].bias]
        return pre, post, bns, params

    def tower_params(self):
        return self._tower_tensors()[3]

    def prepare(self, dev):
        pre, post, bns, params = self._tower_tensors()
        key = (str(dev),) + tuple(p.data_ptr() for p in params) + tuple(
            (p.grad.data_ptr() if p.grad is not None else 0) for p in params) + (
            tuple(b.running_mean.data_ptr() for b in bns) if bns is not None and bns[0].running_mean is not None else ())
        if key == self._key:
            return
        T, Ft, Fo = len(self.convs), self.Ft, self.Fo_t            # real per-tower widths
        Fp, Fq = self.Ftp, self.Fop                                # padded per-tower widths (multiples of 4)
        self.F, self.Fo = T * Fp, T * Fq
        F, FO = self.F, self.Fo
        blocks = post[0].weight.shape[1] // Ft                    # 1 + S * A
        n_pre, n_post = F * 2 * F, FO * blocks * F
        sizes = [n_pre, F, n_post, FO, FO, FO, FO, FO]             # W_pre, b_pre, W_post, b_post, gamma, beta, rm, rv
        offs = np.concatenate([[0], np.cumsum([(s + 3) // 4 * 4 for s in sizes])])
        self.dense = torch.zeros(int(offs[-1]), device=dev)        # zero blocks stay zero: only tower blocks are written
        self.dense_grad = torch.zeros(int(offs[6]), device=dev)
        v = lambda buf, i, shape: buf[int(offs[i]):int(offs[i]) + int(np.prod(shape))].view(shape)
        self.W_pre, self.b_pre = v(self.dense, 0, (F, 2 * F)), v(self.dense, 1, (F,))
        self.W_post, self.b_post = v(self.dense, 2, (FO, blocks * F)), v(self.dense, 3, (FO,))
        self.gamma, self.beta = v(self.dense, 4, (FO,)), v(self.dense, 5, (FO,))
        rm, rv = v(self.dense, 6, (FO,)), v(self.dense, 7, (FO,))
        gviews = [v(self.dense_grad, 0, (F, 2 * F)), v(self.dense_grad, 1, (F,)), v(self.dense_grad, 2, (FO, blocks * F)),
                  v(self.dense_grad, 3, (FO,)), v(self.dense_grad, 4, (FO,)), v(self.dense_grad, 5, (FO,))]
        for t_, g_ in zip((self.W_pre, self.b_pre, self.W_post, self.b_post, self.gamma, self.beta), gviews):
            t_.grad = g_                                           # the fused layer's direct mode accumulates here
        self.tower_bns = bns
        if bns is not None:
            b0 = bns[0]
            self.bn = types.SimpleNamespace(weight=self.gamma, bias=self.beta, running_mean=rm, running_var=rv,
                                            momentum=b0.momentum, eps=b0.eps, track_running_stats=b0.track_running_stats,
                                            num_batches_tracked=None)
            if not b0.track_running_stats or b0.running_mean is None:
                self.bn.running_mean = self.bn.running_var = None
        else:
            self.bn = None
        # staging buffer for the non-direct path: one persistent tensor per tower parameter (fixed addresses)
        self.stage = torch.zeros(sum((p.numel() + 3) // 4 * 4 for p in params), device=dev)
        self.stage_views, o = [], 0
        for p in params:
            self.stage_views.append(self.stage[o:o + p.numel()].view(p.shape))
            o += (p.numel() + 3) // 4 * 4

        def segs(get_a, dense_ptr):
            """segments (tower-side pointer getter, dense-side base pointers) for every tower parameter"""
            rows = []
            per = 6 if bns is not None else 4
            for t in range(T):
                a = [get_a(t * per + i) for i in range(per)]        # (ptr, ld) of W_pre, b_pre, W_post, b_post[, gamma, beta]
                wp, bp, wq, bq = a[0], a[1], a[2], a[3]
                for half in range(2):                                 # [src | dst] halves of the block-diagonal pretrans
                    rows.append((wp[0] + 4 * half * Ft, dense_ptr[0] + 4 * ((t * Fp) * 2 * F + half * F + t * Fp), Ft, Ft,
                                 wp[1], 2 * F))
                rows.append((bp[0], dense_ptr[1] + 4 * t * Fp, 1, Ft, Ft, F))
                for j in range(blocks):                               # h block and every aggregate block
                    rows.append((wq[0] + 4 * j * Ft, dense_ptr[2] + 4 * ((t * Fq) * blocks * F + j * F + t * Fp), Fo, Ft,
                                 wq[1], blocks * F))
                rows.append((bq[0], dense_ptr[3] + 4 * t * Fq, 1, Fo, Fo, FO))
                if bns is not None:
                    rows.append((a[4][0], dense_ptr[4] + 4 * t * Fq, 1, Fo, Fo, FO))
                    rows.append((a[5][0], dense_ptr[5] + 4 * t * Fq, 1, Fo, Fo, FO))
            arr = np.array(rows, dtype=_SEG)
            return torch.from_numpy(arr.view(np.uint8).copy()).to(dev)

        dense_ptrs = [x.data_ptr() for x in (self.W_pre, self.b_pre, self.W_post, self.b_post, self.gamma, self.beta)]
        grad_ptrs = [x.data_ptr() for x in gviews]
        ld = lambda p: p.stride(0) if p.dim() == 2 else p.numel()
        self.t_params = segs(lambda i: (params[i].data_ptr(), ld(params[i])), dense_ptrs)
        self.t_grads_stage = segs(lambda i: (self.stage_views[i].data_ptr(), ld(self.stage_views[i])), grad_ptrs)
        if all(p.grad is not None for p in params):
            self.t_grads_direct = segs(lambda i: (params[i].grad.data_ptr(), ld(params[i].grad)), grad_ptrs)
        else:
            self.t_grads_direct = None
        if self.bn is not None and self.bn.running_mean is not None:
            rows = []
            for t, b in enumerate(bns):
                rows.append((b.running_mean.data_ptr(), rm.data_ptr() + 4 * t * Fq, 1, Fo, Fo, FO))
                rows.append((b.running_var.data_ptr(), rv.data_ptr() + 4 * t * Fq, 1, Fo, Fo, FO))
            self.t_running = torch.from_numpy(np.array(rows, dtype=_SEG).view(np.uint8).copy()).to(dev)
        else:
            self.t_running = torch.zeros(0, dtype=torch.uint8, device=dev)
        self._key = key
        self._serial += 1
        _LIVE.add(self)

    # ---- the layer ------------------------------------------------------------------------------------------------
    def forward(self, g, h, snorm_n):
        t0 = self.convs[0]
        T, N = len(self.convs), h.shape[0]
        if self.Ftp != self.Ft:                                    # zero-pad every tower's column slice (45 -> 48)
            parent = getattr(h, "_dgn_padded", None)
            if (T == 1 and parent is not None and parent.shape == (N, self.Ftp) and parent.data_ptr() == h.data_ptr()
                    and parent._version == h._dgn_padded_version):
                h = parent                                         # the previous padded layer's output: pad columns are 0
            else:
                h = torch.nn.functional.pad(h.reshape(N, T, self.Ft), (0, self.Ftp - self.Ft)).reshape(N, T * self.Ftp)
        out = self._forward_padded(g, h, snorm_n)
        if self.Fop != self.Fo_t:
            if T == 1:
                # a VIEW of the padded rows (leading dimension Fop): the next padded layer takes the parent as it is
                # instead of slicing and re-padding (2 copies forward, 2 backward per layer boundary)
                view = out[:, :self.Fo_t]
                view._dgn_padded, view._dgn_padded_version = out, out._version
                return view
            out = out.reshape(N, T, self.Fop)[:, :, :self.Fo_t].reshape(N, T * self.Fo_t)
        return out

    def _forward_padded(self, g, h, snorm_n):
        t0 = self.convs[0]
        eig = t0._eig(g, h)
        snorm = None
        if t0.graph_norm and snorm_n is not None:
            snorm = snorm_n.reshape(-1)
            if snorm.dtype != torch.float32 or not snorm.is_contiguous():
                snorm = snorm.float().contiguous()
        params = self.tower_params()
        return _TowerFused.apply(self, g, ops._f32c(eig), snorm, t0.training, h, *params)
