"""Batched graph container: destination-major CSR on the device behind a DGL-like surface.

The reference hands its layers a DGL-0.4.2 ``BatchedDGLGraph`` (``dgl.batch`` in
realworld_benchmark/data/molecules.py:229).  The engine needs from it only the node count,
the edge list in edge-id order, ``ndata['eig']`` and ``batch_num_nodes``.
``BatchedGraph`` provides exactly that surface (``ndata`` / ``edata`` / ``number_of_nodes()`` /
``number_of_edges()`` / ``edges()`` / ``in_degrees()`` / ``batch_num_nodes``) and owns the
kernel-side layout:

* ``in_ptr [N+1]``, ``in_src [E]``, ``in_eid [E]``  in-edge slots grouped by destination; slots of
  one node keep edge-id order (the order of the reference's mailbox);
* ``out_ptr [N+1]``, ``out_slot [E]``                by-source transpose for the backward pass;
* ``log_deg [N]``, ``snorm_n [N,1]``, ``graph_ptr [B+1]``;
* ``ovf_ptr [N+1]``                                  overflow-group offsets of the eigen-field layout (``DgnField``):
  the per-batch table of normalised eigenvector weights that ``field()`` builds once and all layers share.

Everything a step needs is packed into ONE pinned host buffer and moved with ONE
host-to-device copy; the device arrays are views into that single allocation.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib

_ALIGN = 16


def _round_up(x: int, a: int = _ALIGN) -> int:
    return (x + a - 1) // a * a


class _Pack:
    """Lays named arrays out in one byte buffer (16 B aligned segments)."""

    def __init__(self):
        self.entries, self.size = {}, 0

    def add(self, name, shape, dtype):
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        self.entries[name] = (self.size, tuple(int(s) for s in shape), dtype)
        self.size = _round_up(self.size + nbytes)

    def host_views(self, buf: np.ndarray):
        out = {}
        for name, (off, shape, dtype) in self.entries.items():
            n = int(np.prod(shape, dtype=np.int64))
            out[name] = buf[off:off + n * dtype.itemsize].view(dtype).reshape(shape)
        return out

    def device_views(self, blob: torch.Tensor):
        out = {}
        for name, (off, shape, dtype) in self.entries.items():
            n = int(np.prod(shape, dtype=np.int64))
            tdt = {np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64,
                   np.dtype(np.float32): torch.float32}[dtype]
            out[name] = blob[off:off + n * dtype.itemsize].view(tdt).reshape(shape)
        return out


class BatchedGraph:
    """A mini-batch of graphs with contiguous node ranges (block-diagonal adjacency)."""

    _STRUCT = ("in_ptr", "in_src", "in_eid", "out_ptr", "out_slot", "graph_ptr", "src", "dst", "log_deg", "snorm_n",
               "meta", "ovf_ptr")

    def __init__(self, n_nodes, src, dst, batch_num_nodes=None, batch_num_edges=None, ndata=None, edata=None,
                 pin_memory=None, capacity=None, graph_capacity=None):
        """``capacity=(N_cap, E_cap)`` pads the arrays to a fixed size (isolated padding nodes, unused
        padding slots) so that batches of different sizes share one memory layout - the shape a
        captured CUDA graph is replayed with.  ``number_of_nodes()`` then returns ``N_cap`` (the row
        count every node tensor must have); ``n_real_nodes`` / ``n_real_edges`` are the true sizes and
        ``meta`` = ``[n_real_nodes, n_real_edges, n_graphs, 0]`` travels to the device with the batch."""
        src = np.ascontiguousarray(np.asarray(src), dtype=np.int32)
        dst = np.ascontiguousarray(np.asarray(dst), dtype=np.int32)
        assert src.shape == dst.shape and src.ndim == 1
        self.n_real_nodes, self.n_real_edges = int(n_nodes), int(src.shape[0])
        self.padded = capacity is not None
        if self.padded:
            if capacity[0] < self.n_real_nodes or capacity[1] < self.n_real_edges:
                raise ValueError("batch (%d nodes, %d edges) exceeds capacity %r"
                                 % (self.n_real_nodes, self.n_real_edges, tuple(capacity)))
            self._n, self._e = int(capacity[0]), int(capacity[1])
        else:
            self._n, self._e = self.n_real_nodes, self.n_real_edges
        self.batch_num_nodes = [int(x) for x in (batch_num_nodes if batch_num_nodes is not None else [self._n])]
        self.batch_num_edges = [int(x) for x in (batch_num_edges if batch_num_edges is not None else [self._e])]
        assert sum(self.batch_num_nodes) == self.n_real_nodes
        B = len(self.batch_num_nodes)
        # graph_capacity: room for that many graphs in graph_ptr (device-side collation fills batches of varying size)
        self.graph_capacity = max(int(graph_capacity), B) if graph_capacity else B
        self.max_graph_nodes_bound = None          # set it on padded (replayed) batches to enable the tile kernels
        Nr, Er = self.n_real_nodes, self.n_real_edges
        ndata = dict(ndata or {})
        edata = dict(edata or {})

        pack = _Pack()
        N, E = self._n, self._e
        for name, shape, dt in (("in_ptr", (N + 1,), np.int32), ("in_src", (E,), np.int32), ("in_eid", (E,), np.int32),
                                ("out_ptr", (N + 1,), np.int32), ("out_slot", (E,), np.int32),
                                ("graph_ptr", (self.graph_capacity + 1,), np.int32), ("src", (E,), np.int32),
                                ("dst", (E,), np.int32),
                                ("log_deg", (N,), np.float32), ("snorm_n", (N, 1), np.float32),
                                ("meta", (4,), np.int32), ("ovf_ptr", (N + 1,), np.int32)):
            pack.add(name, shape, dt)
        for k, v in ndata.items():
            v = np.asarray(v)
            pack.add("n:" + k, (N,) + v.shape[1:], v.dtype)
        for k, v in edata.items():
            v = np.asarray(v)
            pack.add("e:" + k, (E,) + v.shape[1:], v.dtype)
        self._pack = pack

        if pin_memory is None:
            pin_memory = torch.cuda.is_available()
        self._host_blob = torch.empty(max(pack.size, _ALIGN), dtype=torch.uint8, pin_memory=pin_memory)
        if self.padded:
            self._host_blob.zero_()
        hv = pack.host_views(self._host_blob.numpy())
        hv["src"][:Er] = src
        hv["dst"][:Er] = dst
        hv["graph_ptr"][0] = 0
        np.cumsum(self.batch_num_nodes, out=hv["graph_ptr"][1:B + 1])
        hv["graph_ptr"][B + 1:] = Nr
        sizes = np.asarray(self.batch_num_nodes, dtype=np.float32)
        # collate(): snorm_n = sqrt(1 / n_g) per node  (rb/data/molecules.py:222-224)
        hv["snorm_n"][:Nr, 0] = np.repeat(np.sqrt(np.float32(1.0) / sizes), self.batch_num_nodes)
        hv["meta"][:] = (Nr, Er, B, 0)
        for k, v in ndata.items():
            hv["n:" + k][:Nr] = np.asarray(v)
        for k, v in edata.items():
            hv["e:" + k][:Er] = np.asarray(v)

        def p(a):
            return a.ctypes.data_as(ctypes.c_void_p)
        # the CSR is built over the REAL edges; padding nodes get empty in/out ranges at the end
        _lib.check(_lib.lib.dgn_build_csr_host(Nr, Er, p(hv["src"]), p(hv["dst"]), p(hv["in_ptr"]), p(hv["in_src"]),
                                               p(hv["in_eid"]), p(hv["out_ptr"]), p(hv["out_slot"]), p(hv["log_deg"])),
                   "dgn_build_csr_host")
        if self.padded:
            hv["in_ptr"][Nr + 1:] = Er
            hv["out_ptr"][Nr + 1:] = Er
        n_ovf = _lib.lib.dgn_build_groups_host(N, p(hv["in_ptr"]), p(hv["ovf_ptr"]))
        _lib.check(min(n_ovf, 0), "dgn_build_groups_host")
        # eigen-field capacity in groups: one per node + overflow groups (<= E/4; fixed for padded layouts)
        self.n_groups = N + (E // 4 + 1 if self.padded else n_ovf)
        self._fields = {}
        self._host = hv
        self.device = torch.device("cpu")
        self._bind(pack.device_views(self._host_blob))

    # ------------------------------------------------------------------------------------------
    def _bind(self, views):
        self._t = {k: views[k] for k in self._STRUCT}
        self.ndata = {k[2:]: v for k, v in views.items() if k.startswith("n:")}
        self.edata = {k[2:]: v for k, v in views.items() if k.startswith("e:")}
        self._c_graph = None
        self._fields = {}

    def to(self, device, non_blocking=True):
        """One host-to-device copy of the packed buffer; returns self (like DGLGraph.to in later DGLs)."""
        device = torch.device(device)
        if device.type == "cpu":
            self.device = device
            self._bind(self._pack.device_views(self._host_blob))
            return self
        blob = self._host_blob.to(device, non_blocking=non_blocking)
        self.device = device
        self._blob = blob
        self._bind(self._pack.device_views(blob))
        return self

    def copy_into(self, device_blob: torch.Tensor, non_blocking=True):
        """H2D copy of this batch into an existing device buffer of the same (padded) layout."""
        device_blob.copy_(self._host_blob, non_blocking=non_blocking)

    def bind_device_blob(self, blob: torch.Tensor):
        """Make the device-side views point into ``blob`` (a static buffer replayed by a CUDA graph)."""
        assert blob.numel() == self._host_blob.numel()
        self.device = blob.device
        self._blob = blob
        self._bind(self._pack.device_views(blob))
        return self

    @property
    def h2d_bytes(self) -> int:
        return int(self._pack.size)

    @property
    def n_rows_dev(self):
        """Device pointer to the real node count (None when the batch is not padded)."""
        return self._t["meta"] if (self.padded and self.device.type == "cuda") else None

    def check_overflow(self) -> None:
        """After a device-side collation (``dgn_collate_device``): raises if the selected graphs did not fit the fixed
        capacities (the batch was truncated on the device and flagged in ``meta[3]``).  Synchronises - call it lazily,
        e.g. once per epoch or when a loss looks wrong."""
        if self.padded and self.device.type == "cuda" and int(self._t["meta"][3].item()) != 0:
            from ._lib import DgnError
            raise DgnError("a device-collated batch exceeded the capacity (%d nodes, %d edges, %d graphs)"
                           % (self._n, self.number_of_edges(), self.graph_capacity))

    # ---- DGL-like surface ----------------------------------------------------------------------
    def number_of_nodes(self):
        return self._n

    def number_of_edges(self):
        return self._e

    @property
    def batch_size(self):
        """Number of graph slots the readouts run over (graphs past the real count are empty segments)."""
        return self.graph_capacity if self.batch_num_nodes is None else len(self.batch_num_nodes)

    def edges(self):
        return self._t["src"].long(), self._t["dst"].long()

    def in_degrees(self):
        p = self._t["in_ptr"]
        return (p[1:] - p[:-1]).long()

    # ---- engine-side accessors -----------------------------------------------------------------
    def __getattr__(self, name):
        t = self.__dict__.get("_t")
        if t is not None and name in t:
            return t[name]
        raise AttributeError(name)

    def host(self, name):
        return self._host[name]

    def c_graph(self) -> "_lib.DgnGraph":
        """DgnGraph struct over the DEVICE arrays (cached)."""
        if self.device.type != "cuda":
            raise _lib.DgnError("BatchedGraph is on %s: the aggregation kernels need it on a CUDA device "
                                "(call .to('cuda')); there is no CPU path" % self.device)
        if self._c_graph is None:
            t = self._t
            # graph boundaries (tile kernels for high-degree batches).  A padded batch is replayed with other
            # contents, so the bound on the largest graph has to be given explicitly (max_graph_nodes_bound).
            bound = self.max_graph_nodes_bound
            if bound is None and not self.padded and self.batch_num_nodes:
                bound = max(self.batch_num_nodes)
            self._c_graph = _lib.DgnGraph(self._n, self._e, t["in_ptr"].data_ptr(), t["in_src"].data_ptr(),
                                          t["in_eid"].data_ptr(), t["out_ptr"].data_ptr(), t["out_slot"].data_ptr(),
                                          t["log_deg"].data_ptr(), t["graph_ptr"].data_ptr() if bound else None,
                                          self.batch_size if bound else 0, int(bound or 0))
        return self._c_graph

    # ---- eigen-field (DgnField): per-batch normalised eigenvector weights shared by all layers -------------
    def field(self, spec, eig):
        """The ``DgnField`` of this batch for ``spec``'s directional aggregators, built on first use (one launch of
        ``dgn_field_build``) and shared by every layer / direction that aggregates with the same directional list.
        Rebuilt when ``eig`` was modified in place (sign flips, rb/train/train_molecules_graph_regression.py:29-33)
        or after ``invalidate_fields()``; the device buffers are allocated once per graph object."""
        key = tuple((a.kind, a.eig_idx, a.alpha) for a in spec.aggregators if a.kind >= _lib.AGG_DIR_AV)
        ent = self._fields.get(key)
        if ent is None:
            ns = _lib.lib.dgn_field_slots(ctypes.byref(spec.c))
            _lib.check(min(ns, 0), "dgn_field_slots")
            dev = self._t["ovf_ptr"].device
            groups = torch.empty(self.n_groups * (1 + ns) * 4, device=dev, dtype=torch.float32)
            wsum = torch.empty(max((ns + 3) // 4 * 4 * self._n, 4), device=dev, dtype=torch.float32)
            cf = _lib.DgnField(self.n_groups, ns, self._t["ovf_ptr"].data_ptr(), groups.data_ptr(),
                               wsum.data_ptr() if ns > 0 else None)
            ent = {"c": cf, "groups": groups, "wsum": wsum, "stamp": None}
            self._fields[key] = ent
        # (address, version) identifies the contents only while the tensor is alive: the entry keeps a reference so
        # that the allocator cannot hand the same address to a different eig tensor behind our back
        stamp = (eig.data_ptr(), eig._version, eig.stride(0)) if eig is not None else (0, 0, 0)
        if ent["stamp"] != stamp:
            ent["eig_ref"] = eig
            from . import ops
            _lib.check(_lib.lib.dgn_field_build(ctypes.byref(self.c_graph()), ctypes.byref(spec.c),
                                                eig.data_ptr() if eig is not None else None,
                                                eig.stride(0) if eig is not None else 0, ctypes.byref(ent["c"]),
                                                torch.cuda.current_stream(self.device).cuda_stream), "dgn_field_build")
            ops._count(1)
            ent["stamp"] = stamp
        return ent["c"]

    def invalidate_fields(self):
        """Forget which eigen-fields are current (the batch buffers were overwritten, or a CUDA-graph capture
        starts and must record the build launch); the device buffers are kept."""
        for ent in self._fields.values():
            ent["stamp"] = None
            ent["eig_ref"] = None

    @property
    def max_in_degree(self):
        p = self._host["in_ptr"]
        return int((p[1:] - p[:-1]).max()) if self._n else 0


def collate(samples, node_key="feat", edge_key="feat", extra_ndata=(), capacity=None, graph_capacity=None):
    """``dataset.collate`` + ``dgl.batch`` (rb/data/molecules.py:219-230) for synthetic samples.

    Returns ``(graph, labels)``; ``graph.ndata`` holds ``feat`` and ``eig``, ``graph.edata`` holds
    ``feat``; ``graph.snorm_n`` is the per-node graph-norm factor.
    """
    sizes = [int(s["n"]) for s in samples]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    src = np.concatenate([s["src"].astype(np.int64) + offs[i] for i, s in enumerate(samples)])
    dst = np.concatenate([s["dst"].astype(np.int64) + offs[i] for i, s in enumerate(samples)])
    ndata = {node_key: np.concatenate([np.asarray(s["node_feat"]) for s in samples], 0),
             "eig": np.concatenate([np.asarray(s["eig"], dtype=np.float32) for s in samples], 0)}
    for k in extra_ndata:
        ndata[k] = np.concatenate([np.asarray(s[k]) for s in samples], 0)
    edata = {edge_key: np.concatenate([np.asarray(s["edge_feat"]) for s in samples], 0)}
    g = BatchedGraph(int(offs[-1]), src, dst, sizes, [len(s["src"]) for s in samples], ndata, edata,
                     capacity=capacity, graph_capacity=graph_capacity)
    if np.ndim(samples[0]["label"]) == 0:
        labels = torch.from_numpy(np.asarray([s["label"] for s in samples]))
    else:
        labels = torch.from_numpy(np.concatenate([np.asarray(s["label"]) for s in samples]))
    return g, labels
