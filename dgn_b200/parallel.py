"""Batch-sharded data parallelism: one process per GPU, graphs split across ranks.

Graphs of a mini-batch never exchange messages (block-diagonal batched adjacency,
realworld_benchmark/data/molecules.py:229), so every rank aggregates its own shard with no
data-path collective.  The only exchange of a training step is the gradient all-reduce: all
parameters (and their gradients) are views into ONE flat fp32 buffer, so the step issues a single
NCCL all-reduce of ~0.1-0.5 M floats (latency-bound over NVLink/NVSwitch; bucketing or overlap would
buy nothing at this size) and a single fused optimizer update over the flat buffer.

Semantics versus the single-GPU reference: BatchNorm uses per-rank (per-shard) batch
statistics; a mean loss over the GLOBAL batch is the shard losses weighted by their share of the batch
(``shard_loss_weight``: n_r * world / B, which the 1 / world of the gradient average turns into n_r / B), so ranks with
unequal shards still produce the gradient of the global mean.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def flat_numel(module: torch.nn.Module) -> int:
    """Length of the flat buffers ``flatten_parameters`` builds (every parameter padded to 4 floats)."""
    return sum((p.numel() + 3) // 4 * 4 for p in module.parameters() if p.requires_grad)


def flatten_parameters(module: torch.nn.Module, grad_buffer: torch.Tensor = None):
    """Re-home every parameter (and its .grad) as a view into one flat buffer.

    Returns ``(flat_param, flat_grad)``; ``flat_param.grad is flat_grad`` so an optimizer built on
    ``[flat_param]`` updates the whole model with one kernel.  ``grad_buffer``: use this (e.g. peer-mapped symmetric)
    memory for the gradients instead of allocating.
    """
    params = [p for p in module.parameters() if p.requires_grad]
    if not params:
        raise ValueError("module has no trainable parameters")
    dev, dt = params[0].device, params[0].dtype
    sizes = [(p.numel() + 3) // 4 * 4 for p in params]           # keep every view 16 B aligned
    flat_p = torch.zeros(sum(sizes), device=dev, dtype=dt)
    if grad_buffer is not None:
        if grad_buffer.numel() != flat_p.numel() or grad_buffer.dtype != dt or grad_buffer.device != dev:
            raise ValueError("grad_buffer must be a %s tensor of %d elements on %s" % (dt, flat_p.numel(), dev))
        flat_g = grad_buffer.zero_()
    else:
        flat_g = torch.zeros_like(flat_p)
    off = 0
    for p, sz in zip(params, sizes):
        n = p.numel()
        flat_p[off:off + n].copy_(p.data.reshape(-1))
        p.data = flat_p[off:off + n].view_as(p.data)
        p.grad = flat_g[off:off + n].view_as(p.data)
        off += sz
    flat_p = torch.nn.Parameter(flat_p, requires_grad=True)
    flat_p.grad = flat_g
    # re-point the views at the Parameter's storage (same memory; keeps them alive together)
    return flat_p, flat_g


def shard_samples(samples, rank: int, world: int):
    """Contiguous shard of the graph list for ``rank``, balanced by directed edge count."""
    if world == 1:
        return list(samples)
    edges = torch.tensor([len(s["src"]) for s in samples], dtype=torch.float64)
    target = edges.sum() / world
    bounds, acc, cut = [0], 0.0, 1
    for i, e in enumerate(edges.tolist()):
        acc += e
        if acc >= target * cut and len(bounds) < world and (len(samples) - (i + 1)) >= (world - len(bounds)):
            bounds.append(i + 1)
            cut += 1
    while len(bounds) < world:
        bounds.append(min(bounds[-1] + 1, len(samples)))
    bounds.append(len(samples))
    return list(samples[bounds[rank]:bounds[rank + 1]])


def shard_loss_weight(n_local: int, n_global: int, world: int) -> float:
    """Factor for a rank's MEAN loss over its ``n_local`` units (graphs, or nodes for node-level losses) so that the
    average of the rank gradients is the gradient of the mean over all ``n_global`` units."""
    return float(n_local) * float(world) / float(max(n_global, 1))


def allreduce_sum_(flat_grad: torch.Tensor, group=None) -> None:
    """In-place SUM of the flat gradient over the ranks (no-op without an initialised group); the 1 / world of the
    average is folded into the optimizer kernel (``FlatAdam.grad_scale``) instead of a separate launch."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)


def allreduce_mean_(flat_grad: torch.Tensor, group=None) -> None:
    """In-place average of the flat gradient over the ranks (no-op without an initialised group)."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    flat_grad.div_(world)


class PeerGradients:
    """The ranks' flat gradient buffers in symmetric (NVLink peer-mapped) memory + the flags of ``dgn_allreduce_adam``:
    the step's gradient all-reduce and Adam update become ONE launch of our own kernel inside the captured CUDA graph
    (reduce-scatter in rank order over peer loads -> all-gather + Adam), instead of an NCCL all-reduce and an optimizer
    launch outside it.  Construction is collective; raises when the ranks are not all P2P-connected on one node."""

    def __init__(self, numel: int, device, group=None):
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > _lib.AR_MAX_WORLD:
            raise RuntimeError("dgn_allreduce_adam handles up to %d ranks" % _lib.AR_MAX_WORLD)
        if numel % 4:
            raise ValueError("flat buffers are padded to multiples of 4 floats")
        self.grad = symm.empty(numel, dtype=torch.float32, device=device)
        self.flags = symm.empty(3 * _lib.AR_BLOCKS * self.world, dtype=torch.int32, device=device)
        self.grad.zero_()
        self.flags.zero_()
        torch.cuda.synchronize(device)
        hg = symm.rendezvous(self.grad, group.group_name)
        hf = symm.rendezvous(self.flags, group.group_name)
        self._handles = (hg, hf)                                  # keep the mappings alive
        self.grad_ptrs = torch.tensor([int(p) for p in hg.buffer_ptrs], dtype=torch.int64, device=device)
        self.flag_ptrs = torch.tensor([int(p) for p in hf.buffer_ptrs], dtype=torch.int64, device=device)
        self.epoch = torch.zeros(4, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group)                                       # every pad is zero before anybody signals
        import os
        # DGN_PEER_KEEP_SUM=1: also store the all-reduced SUM in ``self.reduced`` (tests, gradient logging)
        self.reduced = torch.zeros_like(self.grad) if os.environ.get("DGN_PEER_KEEP_SUM", "0") == "1" else None
        self.one_shot = self.world <= int(os.environ.get("DGN_AR_ONESHOT_MAX_WORLD", "2"))
        self._pg = _lib.DgnPeerGroup(self.world, self.rank, self.grad_ptrs.data_ptr(), self.flag_ptrs.data_ptr(),
                                     self.epoch.data_ptr(), self.reduced.data_ptr() if self.reduced is not None else None,
                                     self.world if self.one_shot else 0)

    def timed_out(self) -> bool:
        """True when a barrier of some step gave up waiting for a peer (synchronises)."""
        return bool(int(self.epoch[2].item()))

    def allreduce_adam(self, opt) -> None:
        """``opt``: the engine's FlatAdam whose ``g`` is ``self.grad``."""
        import ctypes
        from . import _lib, ops
        assert opt.g.data_ptr() == self.grad.data_ptr()
        _lib.check(_lib.lib.dgn_allreduce_adam(ctypes.byref(self._pg), opt.p.numel(), opt.p.data_ptr(),
                                               opt.exp_avg.data_ptr(), opt.exp_avg_sq.data_ptr(), opt._hyper_host[0],
                                               opt.betas[0], opt.betas[1], opt.eps, opt._hyper_host[1],
                                               opt.hyper.data_ptr(), opt.state.data_ptr(),
                                               torch.cuda.current_stream(self.grad.device).cuda_stream),
                   "dgn_allreduce_adam")
        ops._count(1)


def make_peer_gradients(numel: int, device, group=None):
    """``PeerGradients`` when every rank can build it, else None on EVERY rank (the engine then uses NCCL).
    ``DGN_PEER_ALLREDUCE=0`` disables it."""
    import os
    import warnings
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    if os.environ.get("DGN_PEER_ALLREDUCE", "1") == "0" or torch.device(device).type != "cuda":
        return None
    peer, err = None, None
    try:
        peer = PeerGradients(numel, device, group)
    except Exception as e:                                         # no P2P / symmetric memory unavailable
        err = e
    ok = torch.tensor([1 if peer is not None else 0], dtype=torch.int32, device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok.item()) == 0:
        if err is not None:
            warnings.warn("peer-memory gradient all-reduce unavailable (%s: %s); using NCCL" % (type(err).__name__, err))
        return None
    return peer
