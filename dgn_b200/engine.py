"""Training-step runner: eager, or the whole step captured once into a CUDA graph.

At the headline configuration (ZINC, 128 graphs, ~3 k nodes / ~6.4 k edges) a DGN step is a chain
of ~100 small kernels: it is launch-latency bound, not bandwidth bound (hard part 1).
``TrainStep(graphed=True)`` therefore records ``zero_grad -> forward -> loss -> backward ->
(all-reduce) -> Adam`` ONCE and replays it per batch:

* every batch is padded to a fixed ``capacity=(N_cap, E_cap)`` (``BatchedGraph(capacity=...)``): padding
  nodes have no edges, the real node count travels with the batch in device memory
  (``graph.meta``) and the fused norm kernels read it there, so BatchNorm statistics and all
  gradients are those of the unpadded batch;
* the packed batch is ONE host buffer; a step is one H2D copy into the static device buffer, one
  graph launch and (optionally) one 4-byte D2H read of the loss.

torch.cuda.graphs / torch.optim are plumbing; the captured kernels are the C-ABI kernels of
``libdgn_b200.so`` plus the library GEMMs.
"""
from __future__ import annotations

import torch

import ctypes as C

from . import _lib, ops
import os

from .parallel import allreduce_sum_, flat_numel, flatten_parameters, make_peer_gradients


class FlatAdam:
    """Adam over the flat parameter buffer: one ``dgn_adam_step`` launch per step, replayable from a CUDA graph
    (the step counter lives on the device).  Same update rule as ``torch.optim.Adam(lr, betas, eps, weight_decay)``
    (rb/main_molecules.py:82)."""

    def __init__(self, flat_p, flat_g, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0):
        self.p, self.g = flat_p, flat_g
        self.betas, self.eps = betas, eps
        self.exp_avg = torch.zeros_like(flat_g)
        self.exp_avg_sq = torch.zeros_like(flat_g)
        self.state = torch.zeros(2, dtype=torch.int32, device=flat_g.device)
        # {lr, weight_decay, grad_scale} live in DEVICE memory: by-value kernel arguments are frozen into a captured
        # CUDA graph, these are read by every replay (ReduceLROnPlateau / min_lr stop, rb/main_molecules.py:89-130)
        self._hyper_host = [float(lr), float(weight_decay), float(grad_scale)]
        self.hyper = torch.tensor(self._hyper_host, dtype=torch.float32, device=flat_g.device)

    lr = property(lambda self: self._hyper_host[0], lambda self, v: self.set_lr(v))
    weight_decay = property(lambda self: self._hyper_host[1], lambda self, v: self._set(1, v))
    grad_scale = property(lambda self: self._hyper_host[2], lambda self, v: self._set(2, v))

    def _set(self, i, v):
        self._hyper_host[i] = float(v)
        self.hyper.copy_(torch.tensor(self._hyper_host, dtype=torch.float32), non_blocking=False)

    def set_lr(self, lr):
        """Takes effect on the next step, captured or not (one 12-byte H2D copy, stream ordered)."""
        self._set(0, lr)

    @property
    def param_groups(self):
        """Just enough of the torch.optim surface for ``lr_scheduler.ReduceLROnPlateau``-style loops to read the lr."""
        return [{"lr": self.lr, "weight_decay": self.weight_decay}]

    def step(self):
        _lib.check(_lib.lib.dgn_adam_step(self.p.numel(), self.p.data_ptr(), self.g.data_ptr(), self.exp_avg.data_ptr(),
                                          self.exp_avg_sq.data_ptr(), self._hyper_host[0], self.betas[0], self.betas[1],
                                          self.eps, self._hyper_host[1], self.hyper.data_ptr(), self.state.data_ptr(),
                                          torch.cuda.current_stream(self.p.device).cuda_stream), "dgn_adam_step")
        ops._count(1)


class TrainStep:
    def __init__(self, net, template_graph, targets_like, lr=1e-3, weight_decay=0.0, graphed=True,
                 node_key="feat", edge_key="feat", warmup_iters=3, loss_weight=None):
        """``template_graph``: a (padded, for graphed=True) host ``BatchedGraph`` defining the batch layout.
        ``loss_weight``: this rank's share factor of the global mean loss (``parallel.shard_loss_weight``); it lives in
        device memory (``set_loss_weight``) so a captured step follows batches whose shards differ in size."""
        import torch.distributed as dist
        self.net = net
        self.dev = next(net.parameters()).device
        self.graphed = graphed
        self.node_key, self.edge_key = node_key, edge_key
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        # more than one rank: the gradients live in NVLink peer-mapped memory and ONE launch of our own kernel
        # (dgn_allreduce_adam) all-reduces them and applies Adam, inside the captured graph; NCCL is the fallback
        self.peer = make_peer_gradients(flat_numel(net), self.dev) if self.world > 1 else None
        self.flat_p, self.flat_g = flatten_parameters(net, self.peer.grad if self.peer is not None else None)
        # gradient average over the ranks = SUM all-reduce + 1 / world folded into the Adam kernel
        self.opt = FlatAdam(self.flat_p, self.flat_g, lr=lr, weight_decay=weight_decay, grad_scale=1.0 / self.world)
        self.loss_weight = None
        if loss_weight is not None:
            self.loss_weight = torch.full((), float(loss_weight), device=self.dev)
        if graphed and not template_graph.padded:
            raise ValueError("graphed=True needs batches padded to a fixed capacity (BatchedGraph(capacity=...))")
        # static device-side batch: the graph object is re-bound onto this buffer once
        self.blob = torch.empty(template_graph._host_blob.numel(), dtype=torch.uint8, device=self.dev)
        template_graph.copy_into(self.blob, non_blocking=False)
        self.g = template_graph.bind_device_blob(self.blob)
        self.targets = torch.zeros(targets_like.shape, dtype=targets_like.dtype, device=self.dev)
        self.targets.copy_(targets_like)
        self.loss = torch.zeros((), device=self.dev)
        # NCCL fallback (no peer memory): with more than one rank the NCCL all-reduce of the flat gradient and the Adam launch stay OUTSIDE the captured
        # graph (2 eager launches per step).  Capturing the collective (DGN_GRAPH_ALLREDUCE=1) was tried on B200 / NCCL
        # 2.28.9 / torch 2.11: the capture of a graph that also forks a side stream hung, so it is opt-in only.
        self.split_update = (graphed and self.world > 1 and self.peer is None and
                             os.environ.get("DGN_GRAPH_ALLREDUCE", "0") != "1")
        self.launches_per_step = 0
        self.cuda_graph = None
        if graphed:
            self._capture(warmup_iters)

    # one optimisation step on whatever currently sits in the static batch buffers
    def _fwd_bwd(self):
        g = self.g
        g.invalidate_fields()                 # new batch in the static buffers: the eigen-field is rebuilt (1 launch)
        self.flat_g.zero_()
        # for THIS step only: in-place parameter gradients, weight-gradient GEMMs as a parallel branch, one
        # multi-tensor BatchNorm-counter bump instead of one per layer; the scope joins the side stream on exit
        with ops.step_scope(self.dev) as scope:
            scores = self.net(g, g.ndata[self.node_key], g.edata[self.edge_key], g.snorm_n, None)
            scope.flush_bn_counters()
            loss = self.net.loss(scores, self.targets)
            (loss * self.loss_weight if self.loss_weight is not None else loss).backward()
        return loss

    def set_loss_weight(self, w: float):
        """This rank's share factor for the NEXT step (stream-ordered 4-byte copy; see ``parallel.shard_loss_weight``)."""
        if self.loss_weight is None:
            raise ValueError("construct TrainStep(loss_weight=...) to train with weighted shards")
        self.loss_weight.fill_(float(w))

    def _reduce_and_update(self):
        if self.peer is not None:             # one launch: NVLink reduce-scatter + all-gather + Adam (grad_scale)
            self.peer.allreduce_adam(self.opt)
            return
        allreduce_sum_(self.flat_g)           # one NCCL all-reduce of the flat gradient (no-op on a single rank)
        self.opt.step()                       # Adam on grad / world (grad_scale)

    def _step_body(self):
        loss = self._fwd_bwd()
        if not self.split_update:
            self._reduce_and_update()
        return loss

    def _capture(self, warmup_iters):
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):                 # lazy initialisation (cuBLAS, autograd) outside capture
            for _ in range(max(1, warmup_iters)):   # >= 1: lazy library initialisation must not be captured
                self._step_body()
                if self.split_update:
                    self._reduce_and_update()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        self.cuda_graph = torch.cuda.CUDAGraph()
        before = ops.LAUNCHES
        with torch.cuda.graph(self.cuda_graph):
            loss = self._step_body()
            self.loss.copy_(loss.detach())
        self.launches_per_step = ops.LAUNCHES - before + (1 if self.split_update else 0)

    # ------------------------------------------------------------------------------------------------
    def load(self, host_graph, host_targets):
        """Stage a batch: one H2D copy of the packed graph + the targets (pinned host memory)."""
        if host_graph._host_blob.numel() != self.blob.numel():
            raise ValueError("batch layout differs from the template (capacity / feature keys must match)")
        self.blob.copy_(host_graph._host_blob, non_blocking=True)
        self.targets.copy_(host_targets, non_blocking=True)

    def load_ids(self, dataset, ids_host):
        """Stage the batch made of the dataset graphs ``ids_host`` (pinned int32 ``[B]``): a ``4 B``-byte H2D copy of the
        index list, then ONE launch assembles the batch (and its targets) in HBM from the dataset-resident fragments
        (``dgn_b200.data.device_dataset.DeviceDataset``) - the host never touches graph data in the training loop."""
        if self.cuda_graph is not None and int(ids_host.numel()) != self.g.graph_capacity:
            raise ValueError("a captured step needs exactly %d graphs per batch (got %d)"
                             % (self.g.graph_capacity, int(ids_host.numel())))
        if getattr(self, "_ids", None) is None or self._ids.numel() != ids_host.numel():
            self._ids = torch.empty(ids_host.numel(), dtype=torch.int32, device=self.dev)
        self._ids.copy_(ids_host, non_blocking=True)
        dataset.collate_into(self.g, self._ids, self.targets if dataset.targets is not None else None)

    def load_device(self, device_blob, device_targets):
        """Stage a batch that is already resident in HBM (device-to-device copy)."""
        self.blob.copy_(device_blob, non_blocking=True)
        self.targets.copy_(device_targets, non_blocking=True)

    def run_logged(self):
        """``run()`` + an asynchronous 4-byte D2H copy of the loss into a pinned ring.  Returns ``(value, event)``:
        ``value`` is a pinned host scalar that holds the loss once ``event.synchronize()`` returns.  A training loop
        that reads the loss of step i after it has queued step i + 1 never leaves the GPU waiting for the host
        (the ring has 4 slots: read a value before 4 more steps are queued)."""
        import torch as _t
        if getattr(self, "_ring", None) is None:
            self._ring = _t.empty(4, dtype=_t.float32).pin_memory()
            self._ring_ev = [_t.cuda.Event() for _ in range(4)]
            self._ring_i = 0
        i = self._ring_i = (self._ring_i + 1) % 4
        self._ring[i:i + 1].copy_(self.run().reshape(1), non_blocking=True)
        self._ring_ev[i].record()
        return self._ring[i], self._ring_ev[i]

    def run(self):
        """One training step on the staged batch; returns the loss as a device scalar."""
        if self.cuda_graph is not None:
            self.cuda_graph.replay()
            if self.split_update:
                self._reduce_and_update()
            return self.loss
        before = ops.LAUNCHES
        loss = self._step_body()
        self.launches_per_step = ops.LAUNCHES - before
        self.loss.copy_(loss.detach())
        return self.loss
