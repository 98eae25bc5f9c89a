"""pyarrow <-> GpuPromRangeExec (C++ plan layer, csrc/b2p_plan.cpp) over the Arrow C Data Interface.

`PromRangeExec` takes the constructor arguments of the reference's plan nodes with their own names
(SeriesDivide tag_columns/time_index, SeriesNormalize offset/need_filter_out_nan, RangeManipulate
start/end/interval/range/field column, the prom_* UDF name, optional by-label aggregate) and is fed
pyarrow RecordBatches exactly like the reference's tests feed a MemoryExec.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

from . import _lib
from .engine import B2PError, Context, make_params


class _ArrowArray(C.Structure):
    _fields_ = [("length", C.c_int64), ("null_count", C.c_int64), ("offset", C.c_int64), ("n_buffers", C.c_int64),
                ("n_children", C.c_int64), ("buffers", C.c_void_p), ("children", C.c_void_p),
                ("dictionary", C.c_void_p), ("release", C.c_void_p), ("private_data", C.c_void_p)]


class _ArrowSchema(C.Structure):
    _fields_ = [("format", C.c_char_p), ("name", C.c_char_p), ("metadata", C.c_char_p), ("flags", C.c_int64),
                ("n_children", C.c_int64), ("children", C.c_void_p), ("dictionary", C.c_void_p),
                ("release", C.c_void_p), ("private_data", C.c_void_p)]


def _cstr_array(items: Sequence[str]):
    arr = (C.c_char_p * max(len(items), 1))()
    for i, s in enumerate(items):
        arr[i] = s.encode()
    return arr


class PromRangeExec:
    def __init__(self, ctx: Context, function: str, start: int, end: int, interval: int, range: int, time_index: str,
                 field_column: str, tag_columns: Sequence[str], offset: int = 0, need_filter_out_nan: bool = True,
                 param0: float = 0.0, param1: float = 0.0, aggregate: Optional[str] = None,
                 by_columns: Sequence[str] = (), lookback_delta: Optional[int] = None,
                 histogram_quantile: Optional[float] = None, le_column: str = "le"):
        self._L = _lib.load()
        self._ctx = ctx
        p = make_params(0, start, end, interval, range, offset=offset, filter_nan=need_filter_out_nan, param0=param0,
                        param1=param1)
        tags, by = _cstr_array(tag_columns), _cstr_array(by_columns)
        self._h = self._L.b2p_plan_range_create(ctx._h, function.encode(), C.byref(p), time_index.encode(),
                                                field_column.encode(), tags, len(tag_columns),
                                                (aggregate or "").encode(), by, len(by_columns))
        if not self._h:
            raise B2PError(-1, self._L.b2p_plan_last_error().decode())
        if lookback_delta is not None:      # instant-vector selector (InstantManipulate) instead of a range function
            self._L.b2p_plan_set_instant(self._h, int(lookback_delta))
        if histogram_quantile is not None:  # HistogramFold on top
            rc = self._L.b2p_plan_set_histogram_quantile(self._h, le_column.encode(), float(histogram_quantile))
            if rc != 0:
                raise B2PError(rc, self._L.b2p_plan_last_error().decode())

    def push(self, batch) -> None:
        """Feed one pyarrow.RecordBatch (moved into the plan through the C Data Interface)."""
        arr, sch = _ArrowArray(), _ArrowSchema()
        batch._export_to_c(C.addressof(arr), C.addressof(sch))
        rc = self._L.b2p_plan_push_batch(self._h, C.addressof(arr), C.addressof(sch))
        if rc != 0:
            raise B2PError(rc, self._L.b2p_plan_last_error().decode())

    def execute(self):
        """-> pyarrow.RecordBatch with the rows the reference's Filter / Aggregate+Sort would emit."""
        import pyarrow as pa
        arr, sch = _ArrowArray(), _ArrowSchema()
        rc = self._L.b2p_plan_execute(self._h, C.addressof(arr), C.addressof(sch))
        if rc != 0:
            raise B2PError(rc, self._L.b2p_plan_last_error().decode())
        return pa.RecordBatch._import_from_c(C.addressof(arr), C.addressof(sch))

    def num_series(self) -> int:
        return int(self._L.b2p_plan_num_series(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._L.b2p_plan_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass
