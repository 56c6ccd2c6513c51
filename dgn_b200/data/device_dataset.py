"""Dataset-resident graph fragments and device-side collation.

The reference collates every mini-batch on the host (``dgl.batch`` + ``snorm_n`` in ``MoleculeDataset.collate``,
realworld_benchmark/data/molecules.py:219-230) and moves it to the GPU per step.  With 180 GB of HBM the whole dataset
fits on the device many times over, so here it is pre-batched ONCE (``DeviceDataset``: the CSR of all graphs as one
block-diagonal batch, node / edge payloads, per-graph targets) and a training step's host input shrinks to the sampler's
index list: ``dgn_collate_device`` copies the selected graphs' fragments behind each other into the fixed-capacity batch
layout of ``BatchedGraph`` in one launch.  The result is bit-identical to the host ``collate`` of the same graphs.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from dgn_b200 import _lib
from dgn_b200.graph import BatchedGraph, collate


class DeviceDataset:
    def __init__(self, samples, device, node_key="feat", edge_key="feat", targets=None):
        """``samples``: list of graph dicts (dgn_b200.data.synthetic); ``targets``: optional per-graph tensor ``[G, ...]``
        (float32 / int64) that travels with the batch as the step's targets."""
        self.device = torch.device(device)
        self.node_key, self.edge_key = node_key, edge_key
        G = len(samples)
        g, _ = collate(samples, node_key=node_key, edge_key=edge_key)        # the whole dataset as one host batch
        sizes = np.asarray([int(s["n"]) for s in samples], np.int64)
        esizes = np.asarray([len(s["src"]) for s in samples], np.int64)
        self.sizes, self.esizes = sizes, esizes
        node_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        edge_off = np.concatenate([[0], np.cumsum(esizes)]).astype(np.int32)
        ovf = g.host("ovf_ptr")
        ovf_off = ovf[node_off].astype(np.int32)
        dev = self.device
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self._keep = {k: t(g.host(k)) for k in ("in_ptr", "in_src", "in_eid", "out_ptr", "out_slot", "src", "dst",
                                                  "log_deg", "ovf_ptr")}
        self._keep.update(node_off=t(node_off), edge_off=t(edge_off), ovf_off=t(ovf_off))
        self.ndata = {k: v.to(dev).contiguous() for k, v in g.ndata.items()}
        self.edata = {k: v.to(dev).contiguous() for k, v in g.edata.items()}
        self.targets = None if targets is None else targets.to(dev).contiguous()
        k = self._keep
        self.c = _lib.DgnDataset(G, *(k[n].data_ptr() for n in ("node_off", "edge_off", "ovf_off", "in_ptr", "in_src", "in_eid",
                                                                 "out_ptr", "out_slot", "src", "dst", "log_deg", "ovf_ptr")))
        self.n_graphs = G
        self._template_sample = samples[0]

    def precompute_eig(self, k, norm="none", key="eig"):
        """Laplacian eigenvectors of every graph of the dataset ON THE DEVICE (``dgn_eig_precompute``; the reference does
        this per graph on the host with ARPACK, rb/data/molecules.py:100-116): fills ``ndata[key]`` ``[N_total, k]`` and
        returns the eigenvalues ``[G, k]``."""
        code = {"none": 0, "sym": 1, "walk": 2}[norm]
        Nt = int(self._keep["node_off"][-1])
        eig = torch.zeros((Nt, k), device=self.device, dtype=torch.float32)
        val = torch.zeros((self.n_graphs, k), device=self.device, dtype=torch.float32)
        kk = self._keep
        _lib.check(_lib.lib.dgn_eig_precompute(self.n_graphs, kk["node_off"].data_ptr(), kk["in_ptr"].data_ptr(),
                                               kk["in_src"].data_ptr(), int(self.sizes.max()) if self.n_graphs else 0, code,
                                               k, eig.data_ptr(), eig.stride(0), val.data_ptr(),
                                               torch.cuda.current_stream(self.device).cuda_stream), "dgn_eig_precompute")
        from dgn_b200 import ops
        ops._count(1)
        self.ndata[key] = eig
        return val

    def capacity_for(self, batch_size, slack=1.03, quantile_batches=64, seed=0):
        """A (N_cap, E_cap) that holds random batches of ``batch_size`` graphs: the largest of ``quantile_batches``
        sampled batches plus ``slack``, rounded up to multiples of 64."""
        rng = np.random.default_rng(seed)
        n = e = 0
        for _ in range(quantile_batches):
            ids = rng.integers(0, self.n_graphs, size=batch_size)
            n, e = max(n, int(self.sizes[ids].sum())), max(e, int(self.esizes[ids].sum()))
        n, e = int(n * slack) + 64, int(e * slack) + 64
        return (n + 63) // 64 * 64, (e + 63) // 64 * 64

    def template(self, batch_size, capacity):
        """A padded host ``BatchedGraph`` with this dataset's payload keys: defines the static batch layout of a
        ``TrainStep`` (its contents are irrelevant, ``collate_into`` overwrites them)."""
        s = self._template_sample
        g, _ = collate([s], node_key=self.node_key, edge_key=self.edge_key, capacity=capacity,
                       graph_capacity=batch_size)
        g.batch_num_nodes = None                 # batch_size = graph_capacity: all graph slots are live for the readouts
        return g

    def collate_into(self, graph: BatchedGraph, ids_dev: torch.Tensor, targets_out: torch.Tensor = None):
        """One launch: the graphs ``ids_dev`` (device int32) become the contents of the device-bound padded ``graph``
        (and their targets the contents of ``targets_out``)."""
        if not graph.padded or graph.device.type != "cuda":
            raise _lib.DgnError("collate_into needs a padded BatchedGraph bound to device memory")
        o = _lib.DgnBatchOut()
        o.n_cap, o.e_cap, o.b_cap = graph.number_of_nodes(), graph.number_of_edges(), graph.graph_capacity
        for name in ("in_ptr", "in_src", "in_eid", "out_ptr", "out_slot", "src", "dst", "graph_ptr", "ovf_ptr", "meta",
                     "log_deg", "snorm_n"):
            setattr(o, name, getattr(graph, name).data_ptr())
        pls = []
        for k, v in self.ndata.items():
            pls.append((v, graph.ndata[k], 0))
        for k, v in self.edata.items():
            pls.append((v, graph.edata[k], 1))
        if targets_out is not None:
            pls.append((self.targets, targets_out, 2))
        if len(pls) > _lib.MAX_PAYLOADS:
            raise _lib.DgnError("at most %d payload arrays per batch" % _lib.MAX_PAYLOADS)
        for i, (src, dst, per) in enumerate(pls):
            rb = src[0].numel() * src.element_size() if src.dim() > 0 and src.shape[0] > 0 else src.element_size()
            if dst.dtype != src.dtype or rb % 4:
                raise _lib.DgnError("payload %d: dtype / row size mismatch" % i)
            o.payload[i] = _lib.DgnPayload(src.data_ptr(), dst.data_ptr(), rb, per)
        o.n_payloads = len(pls)
        _lib.check(_lib.lib.dgn_collate_device(C.byref(self.c), ids_dev.data_ptr(), int(ids_dev.numel()), C.byref(o),
                                               torch.cuda.current_stream(self.device).cuda_stream), "dgn_collate_device")
        from dgn_b200 import ops
        ops._count(1)
        graph.batch_num_nodes = None             # host-side list is unknown after a device collate (see graph_ptr)
        graph.invalidate_fields()
