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


def flatten_parameters(module: torch.nn.Module):
    """Re-home every parameter (and its .grad) as a view into one flat buffer.

    Returns ``(flat_param, flat_grad)``; ``flat_param.grad is flat_grad`` so an optimizer built on
    ``[flat_param]`` updates the whole model with one kernel.
    """
    params = [p for p in module.parameters() if p.requires_grad]
    if not params:
        raise ValueError("module has no trainable parameters")
    dev, dt = params[0].device, params[0].dtype
    sizes = [(p.numel() + 3) // 4 * 4 for p in params]           # keep every view 16 B aligned
    flat_p = torch.zeros(sum(sizes), device=dev, dtype=dt)
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
