//! `extern "C"` mirror of `include/b200promql.h` — one item per declaration, same order as the header.
//!
//! Layout contract: `B2pRangeParams` is `struct b2p_range_params` (64 bytes, fields in header order; checked from
//! the C side by `rust-shim/tests/layout.c` and from Python by `tests/test_abi.py::test_params_struct_layout_matches_oracle`).
//! Every other type crossing the boundary is an opaque pointer, a fixed-width integer, `f64`, or the Arrow C Data
//! Interface structs (`arrow::ffi::FFI_ArrowArray` / `FFI_ArrowSchema`, which are `#[repr(C)]` copies of `struct
//! ArrowArray` / `struct ArrowSchema`).
#![allow(non_camel_case_types, clippy::too_many_arguments)]

use std::os::raw::{c_char, c_int, c_void};

use arrow::ffi::{FFI_ArrowArray, FFI_ArrowSchema};

pub const B2P_OK: c_int = 0;
pub const B2P_E_INVALID: c_int = -1;
pub const B2P_E_CUDA: c_int = -2;
pub const B2P_E_UNSORTED: c_int = -3;
pub const B2P_E_NOMEM: c_int = -4;
pub const B2P_E_TOO_LARGE: c_int = -5;
pub const B2P_COMM_ID_BYTES: usize = 128;

/// `enum b2p_fn` — the reference's UDF names in comments (src/query/src/promql/planner.rs:2183-2221).
#[repr(i32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum B2pFn {
    Rate = 0,             // prom_rate      ExtrapolatedRate<true, true>
    Increase = 1,         // prom_increase  ExtrapolatedRate<true, false>
    Delta = 2,            // prom_delta     ExtrapolatedRate<false, false>
    Irate = 3,            // prom_irate
    Idelta = 4,           // prom_idelta
    Resets = 5,           // prom_resets
    Changes = 6,          // prom_changes
    CountOverTime = 7,    // prom_count_over_time
    SumOverTime = 8,      // prom_sum_over_time
    AvgOverTime = 9,      // prom_avg_over_time
    MinOverTime = 10,     // prom_min_over_time
    MaxOverTime = 11,     // prom_max_over_time
    LastOverTime = 12,    // prom_last_over_time
    PresentOverTime = 13, // prom_present_over_time
    AbsentOverTime = 14,  // prom_absent_over_time
    StdvarOverTime = 15,  // prom_stdvar_over_time
    StddevOverTime = 16,  // prom_stddev_over_time
    Deriv = 17,           // prom_deriv
    PredictLinear = 18,   // prom_predict_linear      param0 = t (seconds)
    QuantileOverTime = 19, // prom_quantile_over_time  param0 = phi
    HoltWinters = 20,     // prom_double_exponential_smoothing  param0 = sf, param1 = tf
}

impl B2pFn {
    /// The `prom_*` ScalarUDF name the planner writes into the Projection -> the kernel id.
    pub fn from_udf_name(name: &str) -> Option<Self> {
        Some(match name {
            "prom_rate" => Self::Rate,
            "prom_increase" => Self::Increase,
            "prom_delta" => Self::Delta,
            "prom_irate" => Self::Irate,
            "prom_idelta" => Self::Idelta,
            "prom_resets" => Self::Resets,
            "prom_changes" => Self::Changes,
            "prom_count_over_time" => Self::CountOverTime,
            "prom_sum_over_time" => Self::SumOverTime,
            "prom_avg_over_time" => Self::AvgOverTime,
            "prom_min_over_time" => Self::MinOverTime,
            "prom_max_over_time" => Self::MaxOverTime,
            "prom_last_over_time" => Self::LastOverTime,
            "prom_present_over_time" => Self::PresentOverTime,
            "prom_absent_over_time" => Self::AbsentOverTime,
            "prom_stdvar_over_time" => Self::StdvarOverTime,
            "prom_stddev_over_time" => Self::StddevOverTime,
            "prom_deriv" => Self::Deriv,
            "prom_predict_linear" => Self::PredictLinear,
            "prom_quantile_over_time" => Self::QuantileOverTime,
            "prom_holt_winters" | "prom_double_exponential_smoothing" => Self::HoltWinters,
            _ => return None,
        })
    }
}

/// `enum b2p_agg` — aggregators of `create_aggregate_exprs` (planner.rs:2808-2897).
#[repr(i32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum B2pAgg {
    Sum = 0,
    Avg = 1,
    Count = 2,
    Min = 3,
    Max = 4,
    Stddev = 5,
    Stdvar = 6,
}

/// `struct b2p_range_params`: RangeManipulate::new(start, end, interval, range, ..) (range_manipulate.rs:86-110),
/// SeriesNormalize::new(offset, .., need_filter_out_nan, ..) (normalize.rs:66-83) and the UDF scalars.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct B2pRangeParams {
    pub fn_id: i32,
    pub filter_nan: i32,
    pub start: i64,
    pub end: i64,
    pub interval: i64,
    pub range: i64,
    pub offset: i64,
    pub param0: f64,
    pub param1: f64,
}
const _: () = assert!(std::mem::size_of::<B2pRangeParams>() == 64);
const _: () = assert!(std::mem::align_of::<B2pRangeParams>() == 8);

#[repr(C)]
pub struct b2p_ctx {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct b2p_group_index {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct b2p_plan {
    _opaque: [u8; 0],
}

#[link(name = "b200promql")]
extern "C" {
    // ---- context ------------------------------------------------------------------------------------------
    pub fn b2p_create(device: c_int) -> *mut b2p_ctx;
    pub fn b2p_destroy(ctx: *mut b2p_ctx);
    pub fn b2p_last_error() -> *const c_char;
    pub fn b2p_version() -> *const c_char;
    pub fn b2p_set_stream(ctx: *mut b2p_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn b2p_use_own_stream(ctx: *mut b2p_ctx) -> c_int;
    pub fn b2p_sync(ctx: *mut b2p_ctx) -> c_int;
    pub fn b2p_num_steps(start: i64, end: i64, interval: i64) -> i64;
    pub fn b2p_last_slow_series(ctx: *mut b2p_ctx) -> i64;
    pub fn b2p_last_h2d_bytes(ctx: *mut b2p_ctx) -> i64;
    pub fn b2p_last_warp_tier_series(ctx: *mut b2p_ctx) -> i64;
    pub fn b2p_last_kernel_ms(ctx: *mut b2p_ctx, stage: c_int) -> f64;
    pub fn b2p_launch_count(ctx: *mut b2p_ctx) -> i64;

    // ---- device-pointer API (asynchronous on the context's stream) -----------------------------------------
    pub fn b2p_series_offsets_dev(ctx: *mut b2p_ctx, sid: *const u32, n_rows: u64, n_series: u32, offsets: *mut u64) -> c_int;
    pub fn b2p_range_eval_dev(
        ctx: *mut b2p_ctx, p: *const B2pRangeParams, ts: *const i64, val: *const f64, offsets: *const u64,
        n_rows: u64, n_series: u32, out: *mut f64, valid_words: *mut u32,
    ) -> c_int;
    pub fn b2p_range_udf_dev(
        ctx: *mut b2p_ctx, fn_id: i32, ts: *const i64, val: *const f64, n_rows: u64, packed_ranges: *const i64,
        eval_ts: *const i64, n_win: u64, range_length: i64, param0: f64, param1: f64, out: *mut f64, valid: *mut u8,
    ) -> c_int;
    pub fn b2p_instant_select_dev(
        ctx: *mut b2p_ctx, start: i64, end: i64, interval: i64, lookback: i64, offset: i64, ts: *const i64,
        val: *const f64, offsets: *const u64, n_rows: u64, n_series: u32, out: *mut f64, valid_words: *mut u32,
    ) -> c_int;
    pub fn b2p_group_aggregate_dev(
        ctx: *mut b2p_ctx, agg: i32, vals: *const f64, valid_words: *const u32, gid: *const u32, n_series: u32,
        n_groups: u32, t: u64, out_val: *mut f64, out_cnt: *mut u32,
    ) -> c_int;
    pub fn b2p_group_index_create_dev(
        ctx: *mut b2p_ctx, gid: *const u32, n_series: u32, n_groups: u32, out_index: *mut *mut b2p_group_index,
    ) -> c_int;
    pub fn b2p_group_index_destroy(ctx: *mut b2p_ctx, index: *mut b2p_group_index);
    pub fn b2p_group_aggregate_indexed_dev(
        ctx: *mut b2p_ctx, agg: i32, vals: *const f64, valid_words: *const u32, index: *const b2p_group_index, t: u64,
        out_val: *mut f64, out_cnt: *mut u32,
    ) -> c_int;
    pub fn b2p_range_group_sum_indexed_dev(
        ctx: *mut b2p_ctx, p: *const B2pRangeParams, ts: *const i64, val: *const f64, offsets: *const u64, n_rows: u64,
        n_series: u32, index: *const b2p_group_index, g_lo: u32, g_hi: u32, out_sum: *mut f64, out_cnt: *mut u32,
    ) -> c_int;
    pub fn b2p_range_group_sum_fused(ctx: *mut b2p_ctx, p: *const B2pRangeParams, index: *const b2p_group_index) -> c_int;
    pub fn b2p_range_group_sum_dev(
        ctx: *mut b2p_ctx, p: *const B2pRangeParams, ts: *const i64, val: *const f64, offsets: *const u64, n_rows: u64,
        n_series: u32, gid: *const u32, n_groups: u32, out_sum: *mut f64, out_cnt: *mut u32,
    ) -> c_int;
    pub fn b2p_group_aggregate_partial_dev(
        ctx: *mut b2p_ctx, agg: i32, vals: *const f64, valid_words: *const u32, gid: *const u32, n_series: u32,
        n_groups: u32, t: u64, out_val: *mut f64, out_cnt: *mut u32, out_mean: *mut f64,
    ) -> c_int;
    pub fn b2p_group_finalize_dev(ctx: *mut b2p_ctx, agg: i32, val: *mut f64, cnt: *const u32, n: u64) -> c_int;

    // ---- multi-GPU: the library's own NCCL communicator ---------------------------------------------------------
    pub fn b2p_comm_unique_id(out_id: *mut c_void, bytes: usize) -> c_int;
    pub fn b2p_comm_init(ctx: *mut b2p_ctx, id: *const c_void, bytes: usize, n_ranks: c_int, rank: c_int) -> c_int;
    pub fn b2p_comm_destroy(ctx: *mut b2p_ctx) -> c_int;
    pub fn b2p_allreduce_partials_dev(ctx: *mut b2p_ctx, agg: i32, val: *mut f64, cnt: *mut u32, mean: *mut f64, n: u64) -> c_int;
    pub fn b2p_allreduce_columns_dev(ctx: *mut b2p_ctx, sum: *mut f64, cnt: *mut u64, n_cols: u32) -> c_int;
    pub fn b2p_range_group_sum_allreduce_dev(
        ctx: *mut b2p_ctx, p: *const B2pRangeParams, ts: *const i64, val: *const f64, offsets: *const u64, n_rows: u64,
        n_series: u32, index: *const b2p_group_index, n_tiles: i32, out_sum: *mut f64, out_cnt: *mut u32,
    ) -> c_int;

    // ---- HistogramFold / wide scan --------------------------------------------------------------------------------
    pub fn b2p_histogram_quantile_dev(
        ctx: *mut b2p_ctx, phi: f64, le: *const f64, n_buckets: u32, rates: *const f64, valid_words: *const u32,
        n_hist: u32, t: u64, out: *mut f64, out_valid_words: *mut u32,
    ) -> c_int;
    pub fn b2p_histogram_fold_dev(
        ctx: *mut b2p_ctx, phi: f64, hist_off: *const u32, bucket_series: *const u32, bucket_le: *const f64,
        n_hist: u32, rates: *const f64, valid_words: *const u32, t: u64, out: *mut f64, out_valid_words: *mut u32,
    ) -> c_int;
    pub fn b2p_column_reduce_dev(
        ctx: *mut b2p_ctx, cols: *const *const f64, n_cols: u32, n_rows: u64, out_sum: *mut f64, out_cnt: *mut u64,
    ) -> c_int;

    // ---- host-side helper (no device work): SeriesDivide + cadence scan of one sorted batch ---------------------------
    pub fn b2p_host_scan_series(
        ts: *const i64, sid: *const u32, offsets_in: *const u64, n_rows: u64, n_series: u32, sid_base: u32,
        offsets_out: *mut u64, t0: *mut i64, cadence: *mut i64, all_regular: *mut i32,
    ) -> c_int;

    // ---- host-pointer API (synchronous) ------------------------------------------------------------------------------
    pub fn b2p_range_eval(
        ctx: *mut b2p_ctx, p: *const B2pRangeParams, ts: *const i64, val: *const f64, sid: *const u32,
        offsets_host: *const u64, n_rows: u64, n_series: u32, out: *mut f64, valid_words: *mut u32, out_ts: *mut i64,
    ) -> c_int;
    pub fn b2p_range_udf(
        ctx: *mut b2p_ctx, fn_id: i32, ts: *const i64, val: *const f64, n_rows: u64, packed_ranges: *const i64,
        eval_ts: *const i64, n_win: u64, range_length: i64, param0: f64, param1: f64, out: *mut f64, valid: *mut u8,
    ) -> c_int;
    pub fn b2p_instant_select(
        ctx: *mut b2p_ctx, start: i64, end: i64, interval: i64, lookback: i64, offset: i64, ts: *const i64,
        val: *const f64, sid: *const u32, offsets_host: *const u64, n_rows: u64, n_series: u32, out: *mut f64,
        valid_words: *mut u32,
    ) -> c_int;
    pub fn b2p_group_aggregate(
        ctx: *mut b2p_ctx, agg: i32, vals: *const f64, valid_words: *const u32, gid: *const u32, n_series: u32,
        n_groups: u32, t: u64, out_val: *mut f64, out_cnt: *mut u32,
    ) -> c_int;
    pub fn b2p_histogram_quantile(
        ctx: *mut b2p_ctx, phi: f64, le: *const f64, n_buckets: u32, rates: *const f64, valid_words: *const u32,
        n_hist: u32, t: u64, out: *mut f64, out_valid_words: *mut u32,
    ) -> c_int;
    pub fn b2p_range_histogram_fold(
        ctx: *mut b2p_ctx, p: *const B2pRangeParams, ts: *const i64, val: *const f64, sid: *const u32,
        offsets_host: *const u64, n_rows: u64, n_series: u32, phi: f64, hist_off: *const u32,
        bucket_series: *const u32, bucket_le: *const f64, n_hist: u32, out: *mut f64, out_valid_words: *mut u32,
    ) -> c_int;

    // ---- plan-level API over the Arrow C Data Interface -----------------------------------------------------------------
    pub fn b2p_plan_range_create(
        ctx: *mut b2p_ctx, function: *const c_char, p: *const B2pRangeParams, time_index: *const c_char,
        field_column: *const c_char, tag_columns: *const *const c_char, n_tags: i32, aggregate: *const c_char,
        by_columns: *const *const c_char, n_by: i32,
    ) -> *mut b2p_plan;
    pub fn b2p_plan_set_instant(plan: *mut b2p_plan, lookback_delta: i64) -> c_int;
    pub fn b2p_plan_set_histogram_quantile(plan: *mut b2p_plan, le_column: *const c_char, quantile: f64) -> c_int;
    /// MOVES the batch: on success the release callbacks now belong to the plan.
    pub fn b2p_plan_push_batch(plan: *mut b2p_plan, batch: *mut FFI_ArrowArray, schema: *mut FFI_ArrowSchema) -> c_int;
    pub fn b2p_plan_execute(plan: *mut b2p_plan, out: *mut FFI_ArrowArray, out_schema: *mut FFI_ArrowSchema) -> c_int;
    pub fn b2p_plan_num_series(plan: *mut b2p_plan) -> i64;
    pub fn b2p_plan_destroy(plan: *mut b2p_plan);
    pub fn b2p_plan_last_error() -> *const c_char;

    // ---- bench / test utility ------------------------------------------------------------------------------------------
    pub fn b2p_synth_fill_dev(
        ctx: *mut b2p_ctx, series_begin: u64, n_series: u64, n_samples: u32, t0: i64, scrape_ms: i64, jitter_ms: u32,
        with_resets: i32, seed: u64, ts: *mut i64, val: *mut f64, sid: *mut u32,
    ) -> c_int;
}

/// The library's thread-local error message as an owned string.
pub fn last_error() -> String {
    // SAFETY: b2p_last_error returns a NUL-terminated string owned by the library, valid until the next call on this thread.
    unsafe { std::ffi::CStr::from_ptr(b2p_last_error()).to_string_lossy().into_owned() }
}
pub fn plan_last_error() -> String {
    // SAFETY: as above.
    unsafe { std::ffi::CStr::from_ptr(b2p_plan_last_error()).to_string_lossy().into_owned() }
}
