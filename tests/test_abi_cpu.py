"""CPU: the C-ABI library loads, exports every symbol include/dgn_b200.h declares, and its
host-side entry points and argument validation work without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from dgn_b200 import _lib
from dgn_b200.data.synthetic import make_samples
from dgn_b200.graph import collate

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(REPO, "include", "dgn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dgn_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared_symbols()
    assert "dgn_agg_forward" in names and "dgn_agg_backward" in names and len(names) >= 10
    for n in names:
        assert hasattr(_lib.lib, n), "libdgn_b200.so does not export %s" % n
        assert n in _lib.SIGNATURES, "python binding lacks %s" % n
    assert sorted(_lib.SIGNATURES) == names


def test_abi_version_and_status_strings():
    assert _lib.lib.dgn_abi_version() == _lib.ABI_VERSION >= 4
    assert _lib.lib.dgn_status_string(0) == b"ok"
    assert b"invalid" in _lib.lib.dgn_status_string(-1)


def test_struct_sizes_match_header_layout():
    # 5 int32 + 32 + 32 + 32*4 + 4 (+pad) + float
    assert C.sizeof(_lib.DgnAggSpec) == 20 + 32 + 32 + 128 + 4 + 4
    assert C.sizeof(_lib.DgnGraph) == 8 + 6 * 8 + 8 + 8
    assert C.sizeof(_lib.DgnField) == 8 + 3 * 8


def test_group_builder_and_field_slots():
    # overflow groups of the eigen-field layout: max(0, ceil((D - 4) / 4)) per node
    deg = np.array([0, 1, 4, 5, 8, 9, 51, 3], np.int32)
    in_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    ovf = np.zeros(len(deg) + 1, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    tot = _lib.lib.dgn_build_groups_host(len(deg), p(in_ptr), p(ovf))
    want = np.maximum(0, -(-(deg - 4) // 4))
    assert tot == want.sum() and np.array_equal(ovf, np.concatenate([[0], np.cumsum(want)]))
    assert _lib.lib.dgn_build_groups_host(3, None, p(ovf)) == -1
    # slots: dx and dx-no-abs of one eigenvector share a slot, av has its own, balanced is a single slot
    from dgn_b200.nets.aggregators import AGGREGATORS
    from dgn_b200.nets.scalers import SCALERS
    from dgn_b200.ops import AggSpec

    def slots(names):
        spec = AggSpec([AGGREGATORS[n] for n in names.split()], [SCALERS["identity"]], 1.0, 8, 4)
        return _lib.lib.dgn_field_slots(C.byref(spec.c))
    assert slots("mean max") == 0
    assert slots("mean dir1-dx dir1-dx-no-abs") == 1
    assert slots("dir1-dx dir2-dx dir1-dx-no-abs dir2-dx-no-abs dir1-av dir2-av") == 4
    assert slots("dir1-dx-balanced dir1-0.1 dir1-neg-0.1 dir1-av") == 4
    assert slots("dir1-dx dir1-dx dir1-dx-no-abs") == 2        # a repeated aggregator opens a slot of its own
    assert _lib.lib.dgn_field_slots(None) == -1


@pytest.mark.parametrize("kind,kw", [("zinc", {}), ("cifar", dict(n_min=20, n_max=40)), ("pattern", dict(n_min=20, n_max=40))])
def test_batched_graph_carries_the_field_group_layout(kind, kw):
    """ovf_ptr travels in the packed buffer and the eigen-field capacity covers every batch of a padded layout."""
    samples = make_samples(kind, 4, seed=5, **kw)
    g, _ = collate(samples)
    deg = np.diff(g.host("in_ptr"))
    want = np.maximum(0, -(-(deg - 4) // 4))
    assert np.array_equal(g.host("ovf_ptr"), np.concatenate([[0], np.cumsum(want)]))
    assert g.n_groups == g.number_of_nodes() + int(want.sum())
    cap = (g.number_of_nodes() + 17, g.number_of_edges() + 33)
    gp, _ = collate(samples, capacity=cap)
    ovf = gp.host("ovf_ptr")
    assert ovf.shape[0] == cap[0] + 1 and ovf[-1] == want.sum() and np.all(np.diff(ovf[g.number_of_nodes():]) == 0)
    assert gp.n_groups >= cap[0] + int(want.sum())               # fixed capacity: N_cap + E_cap / 4 + 1 groups


def test_null_arguments_are_rejected_without_a_gpu():
    assert _lib.lib.dgn_agg_forward(None, None, None, None) == -1
    assert _lib.lib.dgn_agg_backward(None, None, None, None, None) == -1
    assert _lib.lib.dgn_norm_forward(None, None) == -1
    assert _lib.lib.dgn_readout_forward(1, None, 4, None, 4, 0, None, 4, None) == -1
    with pytest.raises(_lib.DgnError):
        _lib.check(-2, "x")


def test_csr_builder_matches_numpy():
    rng = np.random.default_rng(0)
    n, e = 50, 400
    src = rng.integers(0, n, e).astype(np.int32)
    dst = rng.integers(0, n, e).astype(np.int32)
    in_ptr = np.zeros(n + 1, np.int32); in_src = np.zeros(e, np.int32); in_eid = np.zeros(e, np.int32)
    out_ptr = np.zeros(n + 1, np.int32); out_slot = np.zeros(e, np.int32); logd = np.zeros(n, np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert _lib.lib.dgn_build_csr_host(n, e, p(src), p(dst), p(in_ptr), p(in_src), p(in_eid), p(out_ptr),
                                       p(out_slot), p(logd)) == 0
    order = np.argsort(dst, kind="stable")                       # mailbox order = edge-id order per destination
    assert np.array_equal(in_eid, order) and np.array_equal(in_src, src[order])
    deg = np.bincount(dst, minlength=n)
    assert np.array_equal(in_ptr, np.concatenate([[0], np.cumsum(deg)]))
    assert np.allclose(logd, np.log(deg + 1.0).astype(np.float32))
    # transpose: every out-edge list holds the slots whose source is that node, ascending
    for u in range(n):
        slots = out_slot[out_ptr[u]:out_ptr[u + 1]]
        assert np.all(in_src[slots] == u) and np.all(np.diff(slots) > 0)
    assert out_ptr[-1] == e
    bad = dst.copy(); bad[3] = n
    assert _lib.lib.dgn_build_csr_host(n, e, p(src), p(bad), p(in_ptr), p(in_src), p(in_eid), p(out_ptr),
                                       p(out_slot), p(logd)) == -1


def test_collate_packs_one_buffer_and_mirrors_dgl_surface():
    samples = make_samples("zinc", 5, seed=3)
    g, labels = collate(samples)
    assert g.number_of_nodes() == sum(s["n"] for s in samples)
    assert g.number_of_edges() == sum(len(s["src"]) for s in samples)
    assert g.batch_num_nodes == [s["n"] for s in samples] and labels.shape == (5,)
    assert g.ndata["eig"].shape == (g.number_of_nodes(), 6) and g.ndata["feat"].dtype.is_floating_point is False
    src, dst = g.edges()
    assert int(src.max()) < g.number_of_nodes()
    assert np.allclose(g.snorm_n[:samples[0]["n"], 0].numpy(), 1.0 / np.sqrt(samples[0]["n"]))
    assert g.graph_ptr.tolist()[-1] == g.number_of_nodes()
    assert g.h2d_bytes % 16 == 0
    with pytest.raises(_lib.DgnError):
        g.c_graph()                                              # CPU graph: no silent fallback


def test_ops_refuse_cpu_tensors():
    import torch
    from dgn_b200.nets.dgn_layer import DGNLayer
    samples = make_samples("zinc", 2, seed=1)
    g, _ = collate(samples)
    layer = DGNLayer(8, 8, 0.0, True, True, "mean dir1-dx", "identity amplification", {"log": 1.0}, "simple",
                     True).model
    with pytest.raises(_lib.DgnError):
        layer(g, torch.randn(g.number_of_nodes(), 8), None, g.snorm_n)


def test_registry_keys_cover_the_reference():
    from dgn_b200.nets.aggregators import AGGREGATORS
    from dgn_b200.nets.scalers import SCALERS
    from oracle.mailbox_ops import AGGREGATORS as REF_AGG, SCALERS as REF_SC
    assert set(REF_AGG) <= set(AGGREGATORS) and set(REF_SC) == set(SCALERS)
    assert AGGREGATORS["dir2-smooth"] is AGGREGATORS["dir2-av"] and "dir4-dx" in AGGREGATORS
    with pytest.raises(KeyError):
        from dgn_b200.nets.dgn_layer import DGNLayer
        DGNLayer(8, 8, 0.0, True, True, "mean bogus", "identity", {"log": 1.0}, "simple", True)


def test_state_dict_keys_match_oracle_layers():
    from dgn_b200.nets.dgn_layer import DGNLayer
    from oracle.directional_layers import DGNLayer as RefLayer
    for tn in ("simple", "complex", "towers"):
        args = (20, 20, 0.0, True, True, "mean max dir1-dx", "identity amplification attenuation", {"log": 1.0}, tn,
                True)
        mine = DGNLayer(*args, towers=4, edge_features=False, edge_dim=0).model.state_dict()
        ref = RefLayer(*args, towers=4, edge_features=False, edge_dim=0).model.state_dict()
        assert list(mine) == list(ref)
        assert all(mine[k].shape == ref[k].shape for k in ref)
