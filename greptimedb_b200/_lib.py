"""ctypes binding of libb200promql.so (the C ABI declared in include/b200promql.h).

There is no CPU fallback: if the shared library is missing this module raises, and if no CUDA
device is present `Context()` raises with the library's own error message.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2P_LIB_PATH") or os.path.join(_HERE, "libb200promql.so")  # override: tuning experiments only

# every symbol include/b200promql.h declares (tests/test_abi.py checks the .so exports them all)
EXPORTED_SYMBOLS = [
    "b2p_create", "b2p_destroy", "b2p_last_error", "b2p_version", "b2p_set_stream", "b2p_use_own_stream", "b2p_sync", "b2p_num_steps",
    "b2p_last_slow_series", "b2p_last_h2d_bytes", "b2p_last_warp_tier_series", "b2p_last_kernel_ms", "b2p_launch_count",
    "b2p_series_offsets_dev", "b2p_range_eval_dev", "b2p_range_udf_dev", "b2p_instant_select_dev",
    "b2p_group_aggregate_dev", "b2p_range_group_sum_dev", "b2p_group_finalize_dev", "b2p_histogram_quantile_dev",
    "b2p_group_index_create_dev", "b2p_group_index_destroy", "b2p_group_aggregate_indexed_dev",
    "b2p_range_group_sum_indexed_dev", "b2p_range_group_sum_fused", "b2p_group_aggregate_partial_dev",
    "b2p_comm_unique_id", "b2p_comm_init", "b2p_comm_destroy", "b2p_allreduce_partials_dev",
    "b2p_range_group_sum_allreduce_dev", "b2p_allreduce_columns_dev", "b2p_histogram_fold_dev", "b2p_range_histogram_fold",
    "b2p_column_reduce_dev", "b2p_host_scan_series", "b2p_range_eval", "b2p_range_udf", "b2p_instant_select", "b2p_group_aggregate",
    "b2p_histogram_quantile", "b2p_synth_fill_dev",
    "b2p_plan_range_create", "b2p_plan_set_instant", "b2p_plan_set_histogram_quantile", "b2p_plan_push_batch", "b2p_plan_execute", "b2p_plan_num_series", "b2p_plan_destroy",
    "b2p_plan_last_error",
]


class RangeParams(C.Structure):
    """struct b2p_range_params"""
    _fields_ = [("fn_id", C.c_int32), ("filter_nan", C.c_int32), ("start", C.c_int64), ("end", C.c_int64),
                ("interval", C.c_int64), ("range", C.c_int64), ("offset", C.c_int64),
                ("param0", C.c_double), ("param1", C.c_double)]


class B200LibraryMissing(ImportError):
    pass


_lib = None


def load() -> C.CDLL:
    """Load libb200promql.so; fail loudly when it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200LibraryMissing(
            f"{LIB_PATH} is missing — build it with greptimedb_b200/csrc/build.sh "
            "(or __graft_entry__.build()).  There is no CPU fallback for the GPU path.")
    L = C.CDLL(LIB_PATH)
    vp, i64, u64, u32, i32, dbl = C.c_void_p, C.c_int64, C.c_uint64, C.c_uint32, C.c_int32, C.c_double
    P = C.POINTER(RangeParams)
    sig = {
        "b2p_create": (vp, [C.c_int]),
        "b2p_destroy": (None, [vp]),
        "b2p_last_error": (C.c_char_p, []),
        "b2p_version": (C.c_char_p, []),
        "b2p_set_stream": (C.c_int, [vp, vp]),
        "b2p_use_own_stream": (C.c_int, [vp]),
        "b2p_sync": (C.c_int, [vp]),
        "b2p_num_steps": (i64, [i64, i64, i64]),
        "b2p_last_slow_series": (i64, [vp]),
        "b2p_last_h2d_bytes": (i64, [vp]),
        "b2p_last_warp_tier_series": (i64, [vp]),
        "b2p_last_kernel_ms": (dbl, [vp, C.c_int]),
        "b2p_launch_count": (i64, [vp]),
        "b2p_series_offsets_dev": (C.c_int, [vp, vp, u64, u32, vp]),
        "b2p_range_eval_dev": (C.c_int, [vp, P, vp, vp, vp, u64, u32, vp, vp]),
        "b2p_range_udf_dev": (C.c_int, [vp, i32, vp, vp, u64, vp, vp, u64, i64, dbl, dbl, vp, vp]),
        "b2p_instant_select_dev": (C.c_int, [vp, i64, i64, i64, i64, i64, vp, vp, vp, u64, u32, vp, vp]),
        "b2p_group_aggregate_dev": (C.c_int, [vp, i32, vp, vp, vp, u32, u32, u64, vp, vp]),
        "b2p_range_group_sum_dev": (C.c_int, [vp, P, vp, vp, vp, u64, u32, vp, u32, vp, vp]),
        "b2p_group_finalize_dev": (C.c_int, [vp, i32, vp, vp, u64]),
        "b2p_group_index_create_dev": (C.c_int, [vp, vp, u32, u32, C.POINTER(vp)]),
        "b2p_group_index_destroy": (None, [vp, vp]),
        "b2p_group_aggregate_indexed_dev": (C.c_int, [vp, i32, vp, vp, vp, u64, vp, vp]),
        "b2p_range_group_sum_indexed_dev": (C.c_int, [vp, P, vp, vp, vp, u64, u32, vp, u32, u32, vp, vp]),
        "b2p_range_group_sum_fused": (C.c_int, [vp, P, vp]),
        "b2p_group_aggregate_partial_dev": (C.c_int, [vp, i32, vp, vp, vp, u32, u32, u64, vp, vp, vp]),
        "b2p_comm_unique_id": (C.c_int, [vp, C.c_size_t]),
        "b2p_comm_init": (C.c_int, [vp, vp, C.c_size_t, C.c_int, C.c_int]),
        "b2p_comm_destroy": (C.c_int, [vp]),
        "b2p_allreduce_partials_dev": (C.c_int, [vp, i32, vp, vp, vp, u64]),
        "b2p_range_group_sum_allreduce_dev": (C.c_int, [vp, P, vp, vp, vp, u64, u32, vp, i32, vp, vp]),
        "b2p_allreduce_columns_dev": (C.c_int, [vp, vp, vp, u32]),
        "b2p_histogram_fold_dev": (C.c_int, [vp, dbl, vp, vp, vp, u32, vp, vp, u64, vp, vp]),
        "b2p_range_histogram_fold": (C.c_int, [vp, P, vp, vp, vp, vp, u64, u32, dbl, vp, vp, vp, u32, vp, vp]),
        "b2p_histogram_quantile_dev": (C.c_int, [vp, dbl, vp, u32, vp, vp, u32, u64, vp, vp]),
        "b2p_column_reduce_dev": (C.c_int, [vp, vp, u32, u64, vp, vp]),
        "b2p_host_scan_series": (C.c_int, [vp, vp, vp, u64, u32, u32, vp, vp, vp, vp]),
        "b2p_range_eval": (C.c_int, [vp, P, vp, vp, vp, vp, u64, u32, vp, vp, vp]),
        "b2p_range_udf": (C.c_int, [vp, i32, vp, vp, u64, vp, vp, u64, i64, dbl, dbl, vp, vp]),
        "b2p_instant_select": (C.c_int, [vp, i64, i64, i64, i64, i64, vp, vp, vp, vp, u64, u32, vp, vp]),
        "b2p_group_aggregate": (C.c_int, [vp, i32, vp, vp, vp, u32, u32, u64, vp, vp]),
        "b2p_histogram_quantile": (C.c_int, [vp, dbl, vp, u32, vp, vp, u32, u64, vp, vp]),
        "b2p_synth_fill_dev": (C.c_int, [vp, u64, u64, u32, i64, i64, u32, i32, u64, vp, vp, vp]),
        "b2p_plan_range_create": (vp, [vp, C.c_char_p, P, C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), i32, C.c_char_p,
                                       C.POINTER(C.c_char_p), i32]),
        "b2p_plan_set_instant": (C.c_int, [vp, i64]),
        "b2p_plan_set_histogram_quantile": (C.c_int, [vp, C.c_char_p, dbl]),
        "b2p_plan_push_batch": (C.c_int, [vp, vp, vp]),
        "b2p_plan_execute": (C.c_int, [vp, vp, vp]),
        "b2p_plan_num_series": (i64, [vp]),
        "b2p_plan_destroy": (None, [vp]),
        "b2p_plan_last_error": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)  # AttributeError here means the .so does not match the header
        f.restype = res
        f.argtypes = args
    _lib = L
    return L
