"""Thin Python handle on the C ABI (include/b200promql.h).

`Context` owns one b2p_ctx (one device, one stream).  Two call families, mirroring the header:
  * host API  — numpy arrays in, numpy arrays out (H2D / kernels / D2H inside the library);
  * device API — torch CUDA tensors (or raw pointers) in place, asynchronous until `sync()`.
torch is used only to hold device memory and the current stream; all arithmetic is in the CUDA
library.  Function names follow the reference's UDF names (prom_rate -> "rate", ...).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from ._lib import RangeParams

FN_IDS = {
    "rate": 0, "increase": 1, "delta": 2, "irate": 3, "idelta": 4, "resets": 5, "changes": 6,
    "count_over_time": 7, "sum_over_time": 8, "avg_over_time": 9, "min_over_time": 10,
    "max_over_time": 11, "last_over_time": 12, "present_over_time": 13, "absent_over_time": 14,
    "stdvar_over_time": 15, "stddev_over_time": 16, "deriv": 17, "predict_linear": 18,
    "quantile_over_time": 19, "holt_winters": 20,
}
AGG_IDS = {"sum": 0, "avg": 1, "count": 2, "min": 3, "max": 4, "stddev": 5, "stdvar": 6}

E_UNSORTED = -3


class B2PError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"b200promql error {code}: {msg}")
        self.code = code


def _ptr(x):
    """numpy array / torch tensor / int / None -> void*"""
    if x is None:
        return None
    if isinstance(x, int):
        return C.c_void_p(x)
    if isinstance(x, np.ndarray):
        return C.c_void_p(x.ctypes.data)
    return C.c_void_p(x.data_ptr())  # torch tensor


def num_steps(start: int, end: int, interval: int) -> int:
    return int(_lib.load().b2p_num_steps(start, end, interval))


def make_params(fn, start, end, interval, range_ms, offset=0, filter_nan=True, param0=0.0, param1=0.0) -> RangeParams:
    fid = FN_IDS[fn] if isinstance(fn, str) else int(fn)
    return RangeParams(fid, int(bool(filter_nan)), int(start), int(end), int(interval), int(range_ms), int(offset),
                       float(param0), float(param1))


def pack_ranges(ranges) -> np.ndarray:
    """[(offset, len)] -> RangeArray keys, offset | len<<32 (range_array.rs:247-254)."""
    r = np.asarray(ranges, dtype=np.uint64).reshape(-1, 2)
    return (r[:, 0] | (r[:, 1] << np.uint64(32))).astype(np.int64)


def valid_to_bool(valid_words: np.ndarray, T: int) -> np.ndarray:
    bits = np.unpackbits(np.ascontiguousarray(valid_words).view(np.uint8), axis=1, bitorder="little")
    return bits[:, :T].astype(bool)


class Context:
    def __init__(self, device: int = 0):
        self._L = _lib.load()
        self._h = self._L.b2p_create(int(device))
        if not self._h:
            raise B2PError(-2, self._L.b2p_last_error().decode())
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._L.b2p_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- plumbing ---------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            raise B2PError(rc, self._L.b2p_last_error().decode())

    def set_stream(self, cuda_stream_ptr: Optional[int]):
        """Enqueue on the given cudaStream_t; 0/None is the legacy default stream."""
        self._check(self._L.b2p_set_stream(self._h, C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))

    def use_own_stream(self):
        self._check(self._L.b2p_use_own_stream(self._h))

    def use_torch_stream(self):
        import torch
        self.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def sync(self):
        self._check(self._L.b2p_sync(self._h))

    def last_slow_series(self) -> int:
        return int(self._L.b2p_last_slow_series(self._h))

    def last_h2d_bytes(self) -> int:
        return int(self._L.b2p_last_h2d_bytes(self._h))

    def last_warp_tier_series(self) -> int:
        return int(self._L.b2p_last_warp_tier_series(self._h))

    def kernel_ms(self, stage: int) -> float:
        return float(self._L.b2p_last_kernel_ms(self._h, stage))

    def launch_count(self) -> int:
        return int(self._L.b2p_launch_count(self._h))

    # -- host API (numpy) --------------------------------------------------------------------
    def range_eval(self, p: RangeParams, ts, val, sid=None, offsets=None):
        """-> (out [S,T] f64, valid_words [S,Tw] u32, eval_ts [T] i64)"""
        ts = np.ascontiguousarray(ts, np.int64)
        val = np.ascontiguousarray(val, np.float64)
        if offsets is not None:
            offsets = np.ascontiguousarray(offsets, np.uint64)
            S = offsets.size - 1
        else:
            sid = np.ascontiguousarray(sid, np.uint32)
            S = int(sid.max()) + 1 if sid.size else 0
        return self.range_eval_n(p, ts, val, sid, offsets, S)

    def range_eval_n(self, p, ts, val, sid, offsets, n_series):
        T = num_steps(p.start, p.end, p.interval)
        Tw = (T + 31) // 32
        out = np.zeros((n_series, T), np.float64)
        valid = np.zeros((n_series, Tw), np.uint32)
        ets = np.zeros(T, np.int64)
        self._check(self._L.b2p_range_eval(self._h, C.byref(p), _ptr(ts), _ptr(val), _ptr(sid), _ptr(offsets),
                                           ts.size, n_series, _ptr(out), _ptr(valid), _ptr(ets)))
        return out, valid, ets

    def range_udf(self, fn, ts, val, ranges, eval_ts=None, range_length=0, param0=0.0, param1=0.0):
        """One prom_* UDF call over explicit windows -> (out f64[], valid bool[])."""
        ts = np.ascontiguousarray(ts, np.int64)
        val = np.ascontiguousarray(val, np.float64)
        packed = pack_ranges(ranges)
        n = packed.size
        ets = None if eval_ts is None else np.ascontiguousarray(eval_ts, np.int64)
        out = np.zeros(n, np.float64)
        valid = np.zeros(n, np.uint8)
        fid = FN_IDS[fn] if isinstance(fn, str) else int(fn)
        self._check(self._L.b2p_range_udf(self._h, fid, _ptr(ts), _ptr(val), ts.size, _ptr(packed), _ptr(ets), n,
                                          int(range_length), float(param0), float(param1), _ptr(out), _ptr(valid)))
        return out, valid.astype(bool)

    def instant_select(self, ts, val, start, end, interval, lookback, offset=0, sid=None, offsets=None):
        ts = np.ascontiguousarray(ts, np.int64)
        val = np.ascontiguousarray(val, np.float64)
        if offsets is not None:
            offsets = np.ascontiguousarray(offsets, np.uint64)
            S = offsets.size - 1
        else:
            sid = np.ascontiguousarray(sid, np.uint32)
            S = int(sid.max()) + 1 if sid.size else 0
        T = num_steps(start, end, interval)
        Tw = (T + 31) // 32
        out = np.zeros((S, T), np.float64)
        valid = np.zeros((S, Tw), np.uint32)
        self._check(self._L.b2p_instant_select(self._h, start, end, interval, lookback, offset, _ptr(ts), _ptr(val),
                                               _ptr(sid), _ptr(offsets), ts.size, S, _ptr(out), _ptr(valid)))
        return out, valid

    def group_aggregate(self, agg, vals, valid, gid, n_groups):
        vals = np.ascontiguousarray(vals, np.float64)
        valid = np.ascontiguousarray(valid, np.uint32)
        gid = np.ascontiguousarray(gid, np.uint32)
        S, T = vals.shape
        out = np.zeros((n_groups, T), np.float64)
        cnt = np.zeros((n_groups, T), np.uint32)
        aid = AGG_IDS[agg] if isinstance(agg, str) else int(agg)
        self._check(self._L.b2p_group_aggregate(self._h, aid, _ptr(vals), _ptr(valid), _ptr(gid), S, n_groups, T,
                                                _ptr(out), _ptr(cnt)))
        return out, cnt

    def histogram_quantile(self, phi, le, rates, valid):
        le = np.ascontiguousarray(le, np.float64)
        rates = np.ascontiguousarray(rates, np.float64)
        valid = np.ascontiguousarray(valid, np.uint32)
        B = le.size
        S, T = rates.shape
        H = S // B
        Tw = (T + 31) // 32
        out = np.zeros((H, T), np.float64)
        ov = np.zeros((H, Tw), np.uint32)
        self._check(self._L.b2p_histogram_quantile(self._h, float(phi), _ptr(le), B, _ptr(rates), _ptr(valid), H, T,
                                                   _ptr(out), _ptr(ov)))
        return out, ov

    # -- device API (torch tensors or raw pointers; asynchronous) ----------------------------------
    def series_offsets_dev(self, sid, n_rows, n_series, offsets):
        self._check(self._L.b2p_series_offsets_dev(self._h, _ptr(sid), n_rows, n_series, _ptr(offsets)))

    def range_eval_dev(self, p, ts, val, offsets, n_rows, n_series, out, valid):
        self._check(self._L.b2p_range_eval_dev(self._h, C.byref(p), _ptr(ts), _ptr(val), _ptr(offsets), n_rows,
                                               n_series, _ptr(out), _ptr(valid)))

    def instant_select_dev(self, start, end, interval, lookback, offset, ts, val, offsets, n_rows, n_series, out, valid):
        self._check(self._L.b2p_instant_select_dev(self._h, start, end, interval, lookback, offset, _ptr(ts), _ptr(val),
                                                   _ptr(offsets), n_rows, n_series, _ptr(out), _ptr(valid)))

    def group_aggregate_dev(self, agg, vals, valid, gid, n_series, n_groups, T, out_val, out_cnt):
        aid = AGG_IDS[agg] if isinstance(agg, str) else int(agg)
        self._check(self._L.b2p_group_aggregate_dev(self._h, aid, _ptr(vals), _ptr(valid), _ptr(gid), n_series,
                                                    n_groups, T, _ptr(out_val), _ptr(out_cnt)))

    def range_group_sum_dev(self, p, ts, val, offsets, n_rows, n_series, gid, n_groups, out_sum, out_cnt):
        self._check(self._L.b2p_range_group_sum_dev(self._h, C.byref(p), _ptr(ts), _ptr(val), _ptr(offsets), n_rows,
                                                    n_series, _ptr(gid), n_groups, _ptr(out_sum), _ptr(out_cnt)))

    # group index: an opaque handle (ctypes void pointer) owned by the caller; destroy it with group_index_destroy
    def group_index_create_dev(self, gid, n_series, n_groups):
        h = C.c_void_p()
        self._check(self._L.b2p_group_index_create_dev(self._h, _ptr(gid), n_series, n_groups, C.byref(h)))
        return h

    def group_index_destroy(self, index):
        self._L.b2p_group_index_destroy(self._h, index)

    def group_aggregate_indexed_dev(self, agg, vals, valid, index, T, out_val, out_cnt):
        aid = AGG_IDS[agg] if isinstance(agg, str) else int(agg)
        self._check(self._L.b2p_group_aggregate_indexed_dev(self._h, aid, _ptr(vals), _ptr(valid), index, T,
                                                            _ptr(out_val), _ptr(out_cnt)))

    def range_group_sum_indexed_dev(self, p, ts, val, offsets, n_rows, n_series, index, g_lo, g_hi, out_sum, out_cnt):
        self._check(self._L.b2p_range_group_sum_indexed_dev(self._h, C.byref(p), _ptr(ts), _ptr(val), _ptr(offsets),
                                                            n_rows, n_series, index, g_lo, g_hi, _ptr(out_sum),
                                                            _ptr(out_cnt)))

    def range_group_sum_fused(self, p, index) -> bool:
        return bool(self._L.b2p_range_group_sum_fused(self._h, C.byref(p), index))

    def group_aggregate_partial_dev(self, agg, vals, valid, gid, n_series, n_groups, T, out_val, out_cnt, out_mean=None):
        aid = AGG_IDS[agg] if isinstance(agg, str) else int(agg)
        self._check(self._L.b2p_group_aggregate_partial_dev(self._h, aid, _ptr(vals), _ptr(valid), _ptr(gid), n_series,
                                                            n_groups, T, _ptr(out_val), _ptr(out_cnt), _ptr(out_mean)))

    # multi-GPU: NCCL communicator owned by the context (rank 0 makes the id, every rank calls comm_init)
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._check(self._L.b2p_comm_unique_id(buf, 128))
        return buf.raw

    def comm_init(self, id_bytes: bytes, n_ranks: int, rank: int):
        buf = C.create_string_buffer(bytes(id_bytes), 128)
        self._check(self._L.b2p_comm_init(self._h, buf, 128, n_ranks, rank))

    def comm_destroy(self):
        self._check(self._L.b2p_comm_destroy(self._h))

    def allreduce_partials_dev(self, agg, val, cnt, mean, n):
        aid = AGG_IDS[agg] if isinstance(agg, str) else int(agg)
        self._check(self._L.b2p_allreduce_partials_dev(self._h, aid, _ptr(val), _ptr(cnt), _ptr(mean), n))

    def range_group_sum_allreduce_dev(self, p, ts, val, offsets, n_rows, n_series, index, n_tiles, out_sum, out_cnt):
        self._check(self._L.b2p_range_group_sum_allreduce_dev(self._h, C.byref(p), _ptr(ts), _ptr(val), _ptr(offsets),
                                                              n_rows, n_series, index, n_tiles, _ptr(out_sum),
                                                              _ptr(out_cnt)))

    def allreduce_columns_dev(self, col_sum, col_cnt, n_cols):
        self._check(self._L.b2p_allreduce_columns_dev(self._h, _ptr(col_sum), _ptr(col_cnt), n_cols))

    def group_finalize_dev(self, agg, val, cnt, n):
        aid = AGG_IDS[agg] if isinstance(agg, str) else int(agg)
        self._check(self._L.b2p_group_finalize_dev(self._h, aid, _ptr(val), _ptr(cnt), n))

    def histogram_quantile_dev(self, phi, le, n_buckets, rates, valid, n_hist, T, out, out_valid):
        self._check(self._L.b2p_histogram_quantile_dev(self._h, float(phi), _ptr(le), n_buckets, _ptr(rates),
                                                       _ptr(valid), n_hist, T, _ptr(out), _ptr(out_valid)))

    def histogram_fold_dev(self, phi, hist_off, bucket_series, bucket_le, n_hist, rates, valid, T, out, out_valid):
        self._check(self._L.b2p_histogram_fold_dev(self._h, float(phi), _ptr(hist_off), _ptr(bucket_series),
                                                   _ptr(bucket_le), n_hist, _ptr(rates), _ptr(valid), T, _ptr(out),
                                                   _ptr(out_valid)))

    def column_reduce_dev(self, col_ptrs, n_cols, n_rows, out_sum, out_cnt):
        self._check(self._L.b2p_column_reduce_dev(self._h, _ptr(col_ptrs), n_cols, n_rows, _ptr(out_sum), _ptr(out_cnt)))

    def synth_fill_dev(self, series_begin, n_series, n_samples, t0, scrape_ms, jitter_ms, with_resets, seed, ts, val, sid):
        self._check(self._L.b2p_synth_fill_dev(self._h, series_begin, n_series, n_samples, t0, scrape_ms, jitter_ms,
                                               int(with_resets), seed, _ptr(ts), _ptr(val), _ptr(sid)))
