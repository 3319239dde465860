"""ctypes binding of libpcrcg_b200.so (C ABI declared in include/pcrcg_b200.h).

There is no fallback: if the CUDA library is missing this module raises at import of the first
operator, and every entry point raises ``RuntimeError`` with the library's message on failure.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PCRCG_B200_LIB: another build of the same library (kernel A/B experiments under tools/); never a different implementation
LIB_PATH = os.environ.get("PCRCG_B200_LIB") or os.path.join(_HERE, "libpcrcg_b200.so")

_lib = None

_P = C.c_void_p
_I64 = C.c_int64
_I32 = C.c_int32
_F = C.c_float
_SZ = C.c_size_t

# name -> (restype, argtypes); kept in one table so tests can check it against include/pcrcg_b200.h
SIGNATURES = {
    "pcrcg_last_error": (C.c_char_p, []),
    "pcrcg_version": (C.c_int, []),
    "pcrcg_free": (None, [_P]),
    "pcrcg_profile_enable": (None, [_I32]),
    "pcrcg_profile_classes": (_I32, []),
    "pcrcg_profile_class_name": (C.c_char_p, [_I32]),
    "pcrcg_profile_report": (C.c_int, [_P, _P]),
    "pcrcg_launch_count": (C.c_uint64, []),
    "pcrcg_subsample_ws_bytes": (_SZ, [_I64, _I32]),
    "pcrcg_subsample_batch_dev": (C.c_int, [_P, _I64, _P, _I32, _F, _I32, _P, _P, _P, _SZ, _P]),
    "pcrcg_subsample_batch_host": (C.c_int, [_P, _I64, _P, _I32, _F, _I32, C.POINTER(_P), C.POINTER(_I64), _P]),
    "pcrcg_subsample_ex_ws_bytes": (_SZ, [_I64, _I32, _I32, _I32]),
    "pcrcg_subsample_batch_ex_dev": (C.c_int, [_P, _I64, _P, _I32, _F, _I32, _P, _I32, _P, _I32, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "pcrcg_subsample_batch_ex_host": (C.c_int, [_P, _I64, _P, _I32, _F, _I32, _P, _I32, _P, _I32, C.POINTER(_P), C.POINTER(_I64), _P,
                                                C.POINTER(_P), C.POINTER(_P)]),
    "pcrcg_group_starts_dev": (C.c_int, [_P, _I32, _I32, _P, _P, _P]),
    "pcrcg_radius_ws_bytes": (_SZ, [_I64, _I64, _I32]),
    "pcrcg_radius_build_dev": (C.c_int, [_P, _I64, _P, _I32, _F, _P, _SZ, _P]),
    "pcrcg_radius_query_dev": (C.c_int, [_P, _I64, _P, _I64, _I32, _F, _I32, _I32, _P, _P, _P, _P, _SZ, _P]),
    "pcrcg_radius_query_ws_bytes": (_SZ, [_I64, _I32]),
    "pcrcg_radius_query_cells_dev": (C.c_int, [_P, _I64, _P, _I64, _I32, _F, _I32, _I32, _P, _P, _P, _P, _SZ, _P, _SZ, _I32, _P]),
    "pcrcg_batch_query_host": (C.c_int, [_P, _I64, _P, _I64, _P, _P, _I32, _F, _I32, C.POINTER(_P), C.POINTER(_I32)]),
    "pcrcg_kpconv_ws_bytes": (_SZ, [_I64, _I64, _I32, _I32]),
    "pcrcg_kpconv_forward_dev": (C.c_int, [_P, _I64, _P, _I64, _P, _I32, _I32, _I32, _P, _I32, _P, _I32, _F, _P, _I32, _P, _P, _SZ, _P]),
    "pcrcg_kpconv_forward_split_dev": (C.c_int, [_P, _I64, _P, _I64, _P, _I32, _I32, _I32, _P, _P, _P, _I32, _P, _I32, _P, _I32, _F, _P, _I32, _P, _P, _SZ, _P]),
    "pcrcg_kpconv_forward_stats_dev": (C.c_int, [_P, _I64, _P, _I64, _P, _I32, _I32, _I32, _P, _P, _P, _I32, _P, _I32, _P, _I32, _F, _P, _I32, _P, _P, _SZ, _P, _I32, _P, _P]),
    "pcrcg_gemm_dev": (C.c_int, [_P, _I32, _P, _I32, _I32, _P, _I32, _I32, _I32, _I32, _P, _P]),
    "pcrcg_split_bf16_dev": (C.c_int, [_P, _I32, _I64, _I32, _P, _P, _I32, _P]),
    "pcrcg_gemm_bf16x3_dev": (C.c_int, [_P, _P, _P, _P, _I32, _P, _I32, _I32, _I32, _I32, _P, _P]),
    "pcrcg_gemm_bf16x3_stats_dev": (C.c_int, [_P, _P, _P, _P, _I32, _P, _I32, _I32, _I32, _I32, _P, _P, _I32, _P, _P]),
    "pcrcg_colstats_final_dev": (C.c_int, [_P, _P, _I32, _I32, _F, _P, _P, _P]),
    "pcrcg_set_option": (C.c_int, [C.c_char_p, _I32]),
    "pcrcg_gemm_force_simt": (None, [_I32]),
    "pcrcg_colstats_dev": (C.c_int, [_P, _I64, _I32, _P, _I32, _F, _P, _P, _P]),
    "pcrcg_norm_act_dev": (C.c_int, [_P, _I64, _I32, _P, _I32, _P, _P, _P, _P, _P, _F, _P, _P, _P, _I32, _P, _P]),
    "pcrcg_norm_act_planes_dev": (C.c_int, [_P, _I64, _I32, _P, _I32, _P, _P, _P, _P, _I32, _P, _P, _F, _P, _P, _P, _I32, _P, _P]),
    "pcrcg_max_pool_planes_dev": (C.c_int, [_P, _P, _I64, _I32, _I32, _P, _I32, _I64, _I32, _I32, _P, _P, _I32, _P]),
    "pcrcg_descriptor_head_dev": (C.c_int, [_P, _I64, _I32, _P, _P, _P, _P]),
    "pcrcg_max_pool_dev": (C.c_int, [_P, _I64, _I32, _P, _I32, _I64, _I32, _I32, _P, _P]),
    "pcrcg_projection_ws_bytes": (_SZ, [_I64]),
    "pcrcg_projection_dev": (C.c_int, [_P, _I64, _P, _I32, _I32, _P, _P, _F, _P, _P, _P, _P, _SZ, _P]),
    "pcrcg_project_scatter_dev": (C.c_int, [_P, _I64, _I32, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _F, _P, _P, _P]),
    "pcrcg_project_scatter_batch_dev": (C.c_int, [_P, _I64, _P, _I32, _P, _P, _I32, _I32, _I32, _F, _P, _P, _P]),
    "pcrcg_knn_dev": (C.c_int, [_P, _I64, _P, _I32, _I32, _P, _P]),
    "pcrcg_edge_max_stats_dev": (C.c_int, [_P, _I32, _P, _I32, _P, _I64, _I32, _I32, _P, _I32, _P, _P, _P]),
    "pcrcg_bias_act_dev": (C.c_int, [_P, _I64, _I32, _P, _F, _P, _P]),
    "pcrcg_softmax_rows_dev": (C.c_int, [_P, _I64, _I32, _I32, _F, _P]),
    "pcrcg_l2norm_rows_dev": (C.c_int, [_P, _I64, _I32, _F, _P, _P]),
    "pcrcg_best_match_dev": (C.c_int, [_P, _I64, _P, _I64, _I32, _P, _P, _P]),
    "pcrcg_mutual_dev": (C.c_int, [_P, _P, _I64, _P, _P]),
    "pcrcg_point2node_dev": (C.c_int, [_P, _I64, _P, _P, _P, _I32, _P, _P]),
    "pcrcg_node_counts_dev": (C.c_int, [_P, _P, _I64, _P, _P, _I32, _P, _P, _P]),
    "pcrcg_closest_pool_dev": (C.c_int, [_P, _I64, _I32, _P, _I32, _I64, _I32, _P, _P]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  pcrcg_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError(lib().pcrcg_last_error().decode("utf-8", "replace"))
