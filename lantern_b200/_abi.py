"""ctypes binding of ``include/lantern_b200.h``.

There is no CPU fallback: if the shared library is missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# LANTERN_B200_LIB: path of an alternative build of the same ABI (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("LANTERN_B200_LIB") or os.path.join(_HERE, "liblantern_b200.so")

F32, BF16, F16 = 0, 1, 2
FAMILY_VANILLA, FAMILY_LLAMAGEN, FAMILY_ANOLE, FAMILY_LUMINA = 0, 1, 2, 3
ROW_IMAGE, ROW_NEWLINE, ROW_EOI = 0, 1, 2
OUT_RESIDUAL_TAIL, OUT_UNIFORM_FALLBACK = 1, 2
MAX_SYNTAX = 8
OK, E_INVALID, E_UNSUPPORTED, E_WORKSPACE, E_NO_DEVICE = 0, -1, -2, -3, -4

EXPORTS = (
    "lantern_version", "lantern_last_error", "lantern_accept_workspace_bytes", "lantern_accept_fused",
    "lantern_accept_phases",
    "lantern_sample_tokens", "lantern_kv_compact", "lantern_build_neighbors", "lantern_philox_uniforms",
    "lantern_session_create", "lantern_session_step", "lantern_session_destroy", "lantern_debug_dist_gemm", "lantern_build_neighbors_workspace_bytes", "lantern_call_create", "lantern_call_destroy",
    "lantern_call_uniforms", "lantern_posterior_call", "lantern_build_dynamic_tree", "lantern_draft_sample",
    "lantern_tree_from_candidates", "lantern_session_last_route", "lantern_accept_greedy",
    "lantern_accept_greedy_workspace_bytes",
)


class AcceptCfg(C.Structure):
    _fields_ = [
        ("n_items", C.c_int32), ("n_rows", C.c_int32), ("n_paths", C.c_int32), ("depth", C.c_int32),
        ("vocab", C.c_int32), ("col0", C.c_int32), ("ncols", C.c_int32), ("logits_dtype", C.c_int32),
        ("item_stride", C.c_int64), ("row_stride", C.c_int64),
        ("family", C.c_int32), ("static_tree", C.c_int32),
        ("cfg_scale", C.c_float), ("temperature", C.c_float), ("top_p", C.c_float), ("top_k", C.c_int32),
        ("lantern", C.c_int32), ("lantern_k", C.c_int32), ("lantern_delta", C.c_float),
        ("lantern_delta_m1", C.c_float), ("table_cols", C.c_int32), ("tok_offset", C.c_int32),
        ("n_syntax", C.c_int32), ("syntax_tokens", C.c_int32 * MAX_SYNTAX),
        ("newline_token", C.c_int32), ("eoi_token", C.c_int32), ("retrieve_shared", C.c_int32),
        ("n_uniforms", C.c_int32), ("n_q_rows", C.c_int32), ("bonus_uniform_last", C.c_int32),
        ("philox_seed", C.c_uint64), ("philox_step", C.c_uint64),
    ]


class AcceptIn(C.Structure):
    _fields_ = [
        ("logits_cond", C.c_void_p), ("logits_uncond", C.c_void_p), ("tree_tokens", C.c_void_p),
        ("retrieve", C.c_void_p), ("row_kinds", C.c_void_p), ("nbr_table", C.c_void_p),
        ("uniforms", C.c_void_p), ("node_q", C.c_void_p), ("draft_op", C.c_void_p),
        ("node_qrow", C.c_void_p), ("sib_off", C.c_void_p), ("sib_idx", C.c_void_p),
        ("sib_tokens", C.c_void_p), ("sib_tokens_stride", C.c_int64),
    ]


class AcceptOut(C.Structure):
    _fields_ = [
        ("accept_length", C.c_void_p), ("best_candidate", C.c_void_p), ("token", C.c_void_p),
        ("path_tokens", C.c_void_p), ("select_indices", C.c_void_p), ("n_draws", C.c_void_p),
        ("flags", C.c_void_p), ("sample_p", C.c_void_p),
    ]


class KvCfg(C.Structure):
    _fields_ = [
        ("n_slabs", C.c_int32), ("elem_bytes", C.c_int32), ("n_outer", C.c_int64),
        ("outer_per_batch", C.c_int64), ("n_batch", C.c_int32), ("s_max", C.c_int32),
        ("head_dim", C.c_int32), ("max_keep", C.c_int32),
        ("slab0", C.c_void_p), ("prev_len0", C.c_int32), ("n_keep0", C.c_int32), ("select_i64", C.c_int32),
        ("reserved0", C.c_int32),
    ]


class LanternError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"lantern_b200 error {code}: {message}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Load the C-ABI library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m lantern_b200.build` "
            "(or __graft_entry__.build()); lantern_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    lib.lantern_version.restype = C.c_int
    lib.lantern_last_error.restype = C.c_char_p
    lib.lantern_accept_workspace_bytes.restype = C.c_size_t
    lib.lantern_accept_workspace_bytes.argtypes = [C.POINTER(AcceptCfg)]
    lib.lantern_accept_fused.restype = C.c_int
    lib.lantern_accept_fused.argtypes = [C.POINTER(AcceptCfg), C.POINTER(AcceptIn), C.POINTER(AcceptOut),
                                         C.c_void_p, C.c_size_t, C.c_void_p]
    lib.lantern_accept_phases.restype = C.c_int
    lib.lantern_accept_phases.argtypes = lib.lantern_accept_fused.argtypes + [C.c_int]
    lib.lantern_accept_greedy_workspace_bytes.restype = C.c_size_t
    lib.lantern_accept_greedy_workspace_bytes.argtypes = [C.POINTER(AcceptCfg)]
    lib.lantern_accept_greedy.restype = C.c_int
    lib.lantern_accept_greedy.argtypes = lib.lantern_accept_fused.argtypes
    lib.lantern_sample_tokens.restype = C.c_int
    lib.lantern_sample_tokens.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.c_void_p]
    lib.lantern_kv_compact.restype = C.c_int
    lib.lantern_kv_compact.argtypes = [C.POINTER(KvCfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.lantern_build_neighbors.restype = C.c_int
    lib.lantern_build_neighbors.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                            C.c_size_t, C.c_void_p, C.c_void_p]
    lib.lantern_build_neighbors_workspace_bytes.restype = C.c_size_t
    lib.lantern_build_neighbors_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    lib.lantern_call_create.restype = C.c_int
    lib.lantern_call_create.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
    lib.lantern_call_destroy.restype = None
    lib.lantern_call_destroy.argtypes = [C.c_void_p]
    lib.lantern_call_uniforms.restype = C.c_void_p
    lib.lantern_call_uniforms.argtypes = [C.c_void_p]
    lib.lantern_posterior_call.restype = C.c_int
    lib.lantern_posterior_call.argtypes = [C.c_void_p, C.POINTER(AcceptCfg), C.POINTER(AcceptIn), C.c_void_p, C.c_void_p,
                                           C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.lantern_debug_dist_gemm.restype = C.c_int
    lib.lantern_debug_dist_gemm.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
    lib.lantern_draft_sample.restype = C.c_int
    lib.lantern_draft_sample.argtypes = [C.POINTER(AcceptCfg), C.POINTER(AcceptIn), C.c_int32, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.lantern_build_dynamic_tree.restype = C.c_int
    lib.lantern_build_dynamic_tree.argtypes = [C.c_void_p] * 4 + [C.c_int32] * 7 + [C.c_void_p] * 7
    lib.lantern_tree_from_candidates.restype = C.c_int
    lib.lantern_tree_from_candidates.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                                 C.c_void_p, C.c_void_p]
    lib.lantern_philox_uniforms.restype = None
    lib.lantern_philox_uniforms.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int32, C.c_void_p]
    lib.lantern_session_create.restype = C.c_int
    lib.lantern_session_create.argtypes = [C.POINTER(AcceptCfg), C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
    lib.lantern_session_step.restype = C.c_int
    lib.lantern_session_step.argtypes = [C.c_void_p, C.POINTER(AcceptCfg), C.POINTER(AcceptIn), C.POINTER(AcceptOut)]
    lib.lantern_session_last_route.restype = C.c_int
    lib.lantern_session_last_route.argtypes = [C.c_void_p]
    lib.lantern_session_destroy.restype = None
    lib.lantern_session_destroy.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().lantern_last_error().decode("utf-8", "replace")
        raise LanternError(rc, msg)
