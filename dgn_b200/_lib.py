"""ctypes binding of ``libdgn_b200.so`` (the C ABI declared in ``include/dgn_b200.h``).

There is deliberately no fallback: if the shared library is missing or a symbol cannot be
resolved, importing this module raises, and every op of the package with it.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DGN_LIB_PATH: an alternative build of the same ABI (kernel tuning experiments, tools/ only)
LIB_PATH = os.environ.get("DGN_LIB_PATH") or os.path.join(_HERE, "libdgn_b200.so")

MAX_AGG, MAX_SCALERS, MAX_SLOTS = 32, 4, 8
ABI_VERSION = 13
NORM_WS_PER_COL = 640

# DgnAggKind / DgnScalerKind / DgnMsgMode
AGG_MEAN, AGG_SUM, AGG_MAX, AGG_MIN, AGG_STD, AGG_VAR = 0, 1, 2, 3, 4, 5
AGG_DIR_AV, AGG_DIR_DX, AGG_DIR_DX_NO_ABS, AGG_DIR_DX_BALANCED, AGG_DIR_SOFTMAX = 6, 7, 8, 9, 10
SCALE_IDENTITY, SCALE_AMPLIFICATION, SCALE_ATTENUATION = 0, 1, 2
MSG_SOURCE, MSG_AFFINE, MSG_DENSE = 0, 1, 2
READOUT_SUM, READOUT_MEAN, READOUT_MAX = 0, 1, 2

_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)


class DgnGraph(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("n_edges", C.c_int32), ("in_ptr", C.c_void_p), ("in_src", C.c_void_p),
                ("in_eid", C.c_void_p), ("out_ptr", C.c_void_p), ("out_slot", C.c_void_p), ("log_deg", C.c_void_p),
                ("graph_ptr", C.c_void_p), ("n_graphs", C.c_int32), ("max_graph_nodes", C.c_int32)]


class DgnAggSpec(C.Structure):
    _fields_ = [("n_feat", C.c_int32), ("group_feat", C.c_int32), ("n_eig", C.c_int32), ("n_agg", C.c_int32),
                ("n_scalers", C.c_int32), ("agg_kind", C.c_uint8 * MAX_AGG), ("agg_eig", C.c_uint8 * MAX_AGG),
                ("agg_alpha", C.c_float * MAX_AGG), ("scaler_kind", C.c_uint8 * MAX_SCALERS), ("avg_log", C.c_float)]


class DgnAggIO(C.Structure):
    _fields_ = [("msg_mode", C.c_int32), ("x", C.c_void_p), ("ld_x", C.c_int32), ("q", C.c_void_p),
                ("ld_q", C.c_int32), ("q_bias", C.c_void_p), ("r", C.c_void_p), ("ld_r", C.c_int32), ("h_in", C.c_void_p),
                ("ld_h", C.c_int32), ("eig", C.c_void_p), ("ld_eig", C.c_int32), ("out", C.c_void_p),
                ("ld_out", C.c_int32), ("out_group_stride", C.c_int32), ("h_copy", C.c_void_p),
                ("ld_hcopy", C.c_int32), ("hcopy_group_stride", C.c_int32), ("field", C.c_void_p)]


class DgnField(C.Structure):
    _fields_ = [("n_groups", C.c_int32), ("n_slots", C.c_int32), ("ovf_ptr", C.c_void_p), ("groups", C.c_void_p),
                ("wsum", C.c_void_p)]


class DgnAggGrad(C.Structure):
    _fields_ = [("g_out", C.c_void_p), ("g_hcopy", C.c_void_p), ("d_x", C.c_void_p), ("ld_dx", C.c_int32),
                ("d_q", C.c_void_p), ("ld_dq", C.c_int32), ("d_r", C.c_void_p), ("ld_dr", C.c_int32),
                ("d_h_in", C.c_void_p), ("ld_dh", C.c_int32), ("d_h_addend", C.c_void_p), ("ld_dha", C.c_int32),
                ("edge_ws", C.c_void_p), ("fold_h_in", C.c_int32)]


class DgnNormArgs(C.Structure):
    _fields_ = [("n_rows", C.c_int32), ("n_cols", C.c_int32), ("y", C.c_void_p), ("ld_y", C.c_int32),
                ("y_bias", C.c_void_p), ("snorm", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("running_mean", C.c_void_p),
                ("running_var", C.c_void_p), ("momentum", C.c_float), ("eps", C.c_float), ("training", C.c_int32),
                ("relu", C.c_int32), ("residual", C.c_void_p), ("ld_res", C.c_int32), ("out", C.c_void_p),
                ("ld_o", C.c_int32), ("stats", C.c_void_p), ("n_rows_dev", C.c_void_p), ("stat_parts", C.c_int32)]


class DgnNormGrad(C.Structure):
    _fields_ = [("g_out", C.c_void_p), ("ld_go", C.c_int32), ("d_y", C.c_void_p), ("ld_dy", C.c_int32),
                ("d_residual", C.c_void_p), ("ld_dres", C.c_int32), ("d_gamma", C.c_void_p), ("d_beta", C.c_void_p),
                ("d_bias", C.c_void_p), ("accumulate", C.c_int32), ("scratch", C.c_void_p)]


class DgnHeadArgs(C.Structure):
    _fields_ = [("n_rows", C.c_int32), ("d0", C.c_int32), ("d1", C.c_int32), ("d2", C.c_int32), ("d_out", C.c_int32),
                ("x", C.c_void_p), ("ld_x", C.c_int32), ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p),
                ("b2", C.c_void_p), ("w3", C.c_void_p), ("b3", C.c_void_p), ("a1", C.c_void_p), ("a2", C.c_void_p),
                ("y", C.c_void_p), ("ld_y", C.c_int32)]


class DgnPostArgs(C.Structure):
    _fields_ = [("n_rows", C.c_int32), ("n_lead", C.c_int32), ("n_agg", C.c_int32), ("n_out", C.c_int32),
                ("n_scalers", C.c_int32), ("scaler_kind", C.c_uint8 * MAX_SCALERS), ("avg_log", C.c_float),
                ("log_deg", C.c_void_p), ("cat", C.c_void_p), ("ld_cat", C.c_int32), ("w", C.c_void_p),
                ("ld_w", C.c_int32)]


class DgnPostStats(C.Structure):
    _fields_ = [("stats", C.c_void_p), ("y_bias", C.c_void_p), ("snorm", C.c_void_p), ("n_rows_dev", C.c_void_p)]


MAX_PAYLOADS = 6


class DgnPayload(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("row_bytes", C.c_int32), ("per", C.c_int32)]


class DgnDataset(C.Structure):
    _fields_ = [("n_graphs", C.c_int32), ("node_off", C.c_void_p), ("edge_off", C.c_void_p), ("ovf_off", C.c_void_p),
                ("in_ptr", C.c_void_p), ("in_src", C.c_void_p), ("in_eid", C.c_void_p), ("out_ptr", C.c_void_p),
                ("out_slot", C.c_void_p), ("src", C.c_void_p), ("dst", C.c_void_p), ("log_deg", C.c_void_p),
                ("ovf_ptr", C.c_void_p)]


class DgnBatchOut(C.Structure):
    _fields_ = [("n_cap", C.c_int32), ("e_cap", C.c_int32), ("b_cap", C.c_int32), ("in_ptr", C.c_void_p),
                ("in_src", C.c_void_p), ("in_eid", C.c_void_p), ("out_ptr", C.c_void_p), ("out_slot", C.c_void_p),
                ("src", C.c_void_p), ("dst", C.c_void_p), ("graph_ptr", C.c_void_p), ("ovf_ptr", C.c_void_p),
                ("meta", C.c_void_p), ("log_deg", C.c_void_p), ("snorm_n", C.c_void_p), ("n_payloads", C.c_int32),
                ("payload", DgnPayload * MAX_PAYLOADS)]


class DgnPeerGroup(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("grad_ptrs", C.c_void_p), ("flag_ptrs", C.c_void_p),
                ("epoch", C.c_void_p), ("reduced", C.c_void_p), ("one_shot_max_world", C.c_int32)]


AR_BLOCKS, AR_MAX_WORLD = 148, 8


class DgnHeadGrad(C.Structure):
    _fields_ = [("g_y", C.c_void_p), ("ld_gy", C.c_int32), ("d_x", C.c_void_p), ("ld_dx", C.c_int32),
                ("d_w1", C.c_void_p), ("d_b1", C.c_void_p), ("d_w2", C.c_void_p), ("d_b2", C.c_void_p),
                ("d_w3", C.c_void_p), ("d_b3", C.c_void_p), ("accumulate", C.c_int32)]


# name -> (restype, argtypes); the CPU test-suite checks this table against include/dgn_b200.h
SIGNATURES = {
    "dgn_abi_version": (C.c_int, []),
    "dgn_status_string": (C.c_char_p, [C.c_int]),
    "dgn_last_cuda_error": (C.c_char_p, []),
    "dgn_agg_forward": (C.c_int, [C.POINTER(DgnGraph), C.POINTER(DgnAggSpec), C.POINTER(DgnAggIO), C.c_void_p]),
    "dgn_agg_backward": (C.c_int, [C.POINTER(DgnGraph), C.POINTER(DgnAggSpec), C.POINTER(DgnAggIO),
                                   C.POINTER(DgnAggGrad), C.c_void_p]),
    "dgn_build_csr_host": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dgn_field_slots": (C.c_int, [C.POINTER(DgnAggSpec)]),
    "dgn_build_groups_host": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p]),
    "dgn_field_build": (C.c_int, [C.POINTER(DgnGraph), C.POINTER(DgnAggSpec), C.c_void_p, C.c_int32,
                                  C.POINTER(DgnField), C.c_void_p]),
    "dgn_norm_forward": (C.c_int, [C.POINTER(DgnNormArgs), C.c_void_p]),
    "dgn_norm_pair_forward": (C.c_int, [C.POINTER(DgnNormArgs), C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                        C.c_void_p, C.c_int32, C.c_void_p]),
    "dgn_norm_backward": (C.c_int, [C.POINTER(DgnNormArgs), C.POINTER(DgnNormGrad), C.c_void_p]),
    "dgn_embedding_backward": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                         C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dgn_adam_step": (C.c_int, [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dgn_allreduce_adam": (C.c_int, [C.POINTER(DgnPeerGroup), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                     C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dgn_gemm_ws_floats": (C.c_int64, []),
    "dgn_gemm_tf32x3": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                  C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                  C.c_void_p]),
    "dgn_post_forward": (C.c_int, [C.POINTER(DgnPostArgs), C.c_void_p, C.c_int32, C.POINTER(DgnPostStats),
                                   C.POINTER(C.c_int32), C.c_void_p]),
    "dgn_post_backward": (C.c_int, [C.POINTER(DgnPostArgs), C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "dgn_post_wgrad": (C.c_int, [C.POINTER(DgnPostArgs), C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                 C.c_void_p]),
    "dgn_pre_wgrad": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "dgn_pair_linear_forward": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                          C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "dgn_pair_linear_backward": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                           C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "dgn_pair_gather_backward": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                           C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                           C.c_void_p, C.c_int32, C.c_void_p]),
    "dgn_eig_precompute": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "dgn_eig_flip": (C.c_int, [C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64, C.c_void_p]),
    "dgn_segment_copy": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "dgn_collate_device": (C.c_int, [C.POINTER(DgnDataset), C.c_void_p, C.c_int32, C.POINTER(DgnBatchOut), C.c_void_p]),
    "dgn_head_forward": (C.c_int, [C.POINTER(DgnHeadArgs), C.c_void_p]),
    "dgn_head_backward": (C.c_int, [C.POINTER(DgnHeadArgs), C.POINTER(DgnHeadGrad), C.c_void_p]),
    "dgn_l1_loss_forward": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dgn_l1_loss_backward": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dgn_readout_forward": (C.c_int, [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                      C.c_void_p, C.c_int32, C.c_void_p]),
    "dgn_readout_backward": (C.c_int, [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                       C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                       C.c_int32, C.c_void_p]),
}


class DgnError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "dgn_b200: %s is missing - build it with `make -C dgn_b200/csrc` or "
            "`python -c 'import __graft_entry__ as g; g.build()'`.  There is no CPU/PyTorch fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype, fn.argtypes = res, args
    got = lib.dgn_abi_version()
    if got != ABI_VERSION:
        raise ImportError("dgn_b200: ABI version %d, binding expects %d - rebuild the library" % (got, ABI_VERSION))
    return lib


lib = _load()


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib.dgn_status_string(status).decode()
        if status == -4:
            msg += ": " + lib.dgn_last_cuda_error().decode()
        raise DgnError("%s failed: %s (status %d)" % (what, msg, status))
