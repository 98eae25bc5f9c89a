// b2p_kernel_t.cuh — K2T: thread-per-series tier of the fused range kernel (rate / increase / delta).
//
// The warp-per-series kernel (K2, b2p_kernels.cuh) spends most of its instruction budget on
// finding window edges in parallel and on warp-uniform bookkeeping.  Here ONE THREAD owns one
// series and walks its eval steps sequentially with two cursors — the same two-pointer walk the
// reference does (range_manipulate.rs:730-769) — so per step it costs two short advances, four
// shared-memory reads and the formula; and because the steps are visited in order, the counter
// correction is maintained exactly like ExtrapolatedRate::calc does it: slid by one sample when the
// window slides by one sample, rescanned otherwise (extrapolate_rate.rs:216-238).  Results are
// therefore bit-identical to the reference's own (sliding) code path.
//
// A warp owns 32 consecutive series.  Samples reach the threads through a shared-memory ring
// [kTRing rows][32 series] filled by warp-cooperative, fully coalesced 128-bit loads (8 lanes x 16 B =
// one 128-byte line per series per column); the [row][series] layout makes every per-thread read
// bank-conflict free no matter which row each thread is at.  Timestamps are kept as uint32 ms since
// (start - range) (see K2's TS32 note).
//
// K2T only takes what it can do exactly and fast; everything else is handed to K2 through a work
// list (which in turn hands the rare rest to the slow kernel): a series with a NaN (stale marker), a
// window needing more than ~46 resident rows, the cursor-overshoot quirk, or query
// spans >= 2^31 ms (host-side decision).
#pragma once
#include <cstdint>

#include "b2p_kernels.cuh"

namespace b2p {

#ifndef B2P_TRING
#define B2P_TRING 64
#endif
constexpr int kTRing = B2P_TRING;   // rows of each series resident in shared memory
constexpr int kTChunk = 16;  // rows per series per cooperative load (one 128-byte line per column)

__constant__ double c_rcp_table[kRcpTable];  // RN(1/n), filled by b2p_create

struct ThreadRing {
  uint32_t* ts;  // [kTRing][32]
  double* val;   // [kTRing][32]
  int lane;
  __device__ __forceinline__ uint32_t t(uint32_t i) const { return ts[(i & (kTRing - 1)) * 32 + lane]; }
  __device__ __forceinline__ double v(uint32_t i) const { return val[(i & (kTRing - 1)) * 32 + lane]; }
};

template <int FN>
__global__ void __launch_bounds__(32) range_thread_kernel(const RangeArgs a) {
  using TR = FnTraits<FN>;
  __shared__ uint32_t s_ts[kTRing * 32];
  __shared__ double s_val[kTRing * 32];
  __shared__ unsigned long long s_base[32];  // first row (16-byte aligned pair boundary) of each series
  __shared__ uint32_t s_end[32];             // end index (lead + rows) of each series
  __shared__ uint32_t s_lead[32];
  const int lane = threadIdx.x;
  const ThreadRing ring{s_ts, s_val, lane};
  const int32_t T = (int32_t)a.T;
  const uint32_t interval = (uint32_t)a.interval, range = (uint32_t)a.range;
  const uint32_t n_batches = (a.n_series + 31) / 32;

  for (uint32_t batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
    const uint32_t s = batch * 32 + lane;
    const bool have = s < a.n_series;
    uint64_t row0 = 0, row1 = 0;
    if (have) {
      row0 = a.offsets[s];
      row1 = a.offsets[s + 1];
    }
    const uint32_t n = (uint32_t)(row1 - row0);
    const uint32_t lead = (uint32_t)(row0 & 1ull);
    const uint32_t end_i = lead + n;  // this thread's samples are ring indices [lead, end_i)
    s_base[lane] = row0 - lead;
    s_end[lane] = have ? end_i : 0u;
    s_lead[lane] = lead;
    double* out_s = a.out + (size_t)s * (size_t)T;
    uint32_t* vw_s = a.valid + (size_t)s * a.Tw;

    // ---- RangeManipulate's grid trimming, exact, from the series' first / last timestamps ---------
    int32_t kf = T, kl = -1;
    bool deferred = false;
    if (have && n > 0) {
      const int64_t first_ts = a.ts[row0] + a.offset, last_ts = a.ts[row1 - 1] + a.offset;
      const int64_t rem = rem_euclid(first_ts - a.start, a.interval);
      const int64_t first_aligned = rem == 0 ? first_ts : first_ts + (a.interval - rem);
      const int64_t s2 = a.start > first_aligned ? a.start : first_aligned;
      const int64_t last_aligned = ((last_ts + a.range) / a.interval) * a.interval;
      const int64_t e2 = a.end < last_aligned ? a.end : last_aligned;
      if (s2 <= e2) {
        const int64_t kf64 = (s2 - a.start) / a.interval;
        const int64_t kl64 = floor_div(e2 - a.start, a.interval);
        kf = kf64 < (int64_t)T ? (int32_t)kf64 : T;
        kl = kl64 >= (int64_t)T ? T - 1 : (int32_t)kl64;
      }
    }
    __syncwarp();

    // ---- per-thread sequential state ---------------------------------------------------------------
    uint32_t loaded = 0;            // warp-uniform: ring indices [0, loaded) have been copied for every series
    uint32_t newest = 0;            // rel. timestamp of this series' newest resident sample
    int32_t hi = (int32_t)lead - 1; // last index with ts <= te
    uint32_t lo = lead;             // first index with ts > te - range
    int32_t k = 0;
    double corr = 0.0;              // ExtrapolatedRate's counter_correction
    uint32_t prev_lo = 0xffffffffu, prev_l = 0;  // previous window layout (prev_offset = usize::MAX when unset)
    uint32_t vword = 0;
    // calculate_range cursor state for the overshoot check (range_manipulate.rs:730-769)
    uint32_t rsi = lead, start_delta = 0, last_range_start = lead;

    bool busy = have;  // still has steps to write
    while (__any_sync(0xffffffffu, busy)) {
      // ---- evaluate every step this thread can finalise with the resident rows --------------------
      if (busy && !deferred) {
        const uint32_t avail = loaded < end_i ? loaded : end_i;  // resident indices are [.., avail)
        const bool all_in = loaded >= end_i;
        while (k < T) {
          const uint32_t tlo = (uint32_t)k * interval;
          const uint32_t te = tlo + range;
          double r = 0.0;
          bool ok = false;
          if (k >= kf && k <= kl) {
            if (!(all_in || (avail > lead && newest > te))) break;  // a later sample could still enter: wait
            while (hi + 1 < (int32_t)avail && ring.t((uint32_t)(hi + 1)) <= te) ++hi;
            while ((int32_t)lo <= hi && ring.t(lo) <= tlo) ++lo;
            const uint32_t l = ((int32_t)lo <= hi) ? (uint32_t)(hi + 1 - (int32_t)lo) : 0u;
            // calculate_range's own cursor state (range_manipulate.rs:736-768), to catch quirk C-13: when
            // its cursor starts at or beyond the last sample it reports an empty window whatever is inside
            if (rsi + start_delta >= end_i) {
              if (l > 0) { deferred = true; break; }  // the exact slow path reproduces the reference here
              start_delta = 0;                        // both loops skipped: range_start_index untouched
            } else {
              if (lo < end_i) {  // some sample is newer than the window start: range_start_index follows it
                const bool brk = (uint32_t)(hi + 1) < end_i;
                rsi = (brk && lo > lead) ? lo - 1 : lo;
              }
              if (l > 0) {
                start_delta = lo - last_range_start;
                last_range_start = lo;
              } else {
                start_delta = 0;
              }
            }
            if (l >= 2) {
              const double first_value = ring.v(lo), last_value = ring.v((uint32_t)hi);
              double result_value;
              if constexpr (TR::kCounter) {
                if (prev_lo != 0xffffffffu && lo == prev_lo + 1 && l == prev_l) {  // extrapolate_rate.rs:219-225
                  const double dropped = ring.v(prev_lo);
                  if (first_value < dropped) corr -= dropped;
                  const double before_last = ring.v((uint32_t)hi - 1);
                  if (last_value < before_last) corr += before_last;
                } else {  // :226-233
                  corr = 0.0;
                  double p = first_value;
                  for (uint32_t i = lo + 1; i <= (uint32_t)hi; ++i) {
                    const double c = ring.v(i);
                    if (c < p) corr += p;
                    p = c;
                  }
                }
                result_value = last_value - first_value + corr;
              } else {
                result_value = last_value - first_value;
              }
              prev_lo = lo;
              prev_l = l;
              const double rcp_len = (l - 1 < (uint32_t)kRcpTable) ? c_rcp_table[l - 1] : 0.0;
              r = extrapolate_parts<FN, uint32_t>(result_value, first_value, ring.t(lo), ring.t((uint32_t)hi), l, te,
                                                  range, rcp_len, a.range_secs, a.rcp_rs);
              ok = true;
            } else {
              prev_lo = 0xffffffffu;  // :206-210
            }
          }
          out_s[k] = r;
          if (ok) vword |= 1u << (k & 31);
          ++k;
          if ((k & 31) == 0 || k == T) {
            vw_s[(k - 1) >> 5] = vword;
            vword = 0;
          }
        }
        if (k >= T) busy = false;
      }
      if (deferred) busy = false;

      // ---- cooperative load of the next kTChunk rows of every series that still has rows ------------
      // ring room: a thread must keep rows >= min(lo, prev_lo) - 1 resident
      if (busy) {
        const uint32_t keep = (prev_lo != 0xffffffffu && prev_lo < lo) ? prev_lo : lo;
        if (loaded + kTChunk > keep + (uint32_t)kTRing - 1u && loaded < end_i) {
          deferred = true;  // window (plus lag) longer than the ring: K2 takes this series
          busy = false;
        }
      }
      if (!__any_sync(0xffffffffu, busy)) break;
#pragma unroll 2
      for (int g = 0; g < 8; ++g) {
        const int sl = g * 4 + (lane >> 3);  // series slot this lane loads for
        const uint32_t i0 = loaded + 2u * (uint32_t)(lane & 7);
        const uint32_t endv = s_end[sl], leadv = s_lead[sl];
        const bool in0 = i0 >= leadv && i0 < endv;
        const bool in1 = i0 + 1 >= leadv && i0 + 1 < endv;
        bool bad = false;
        if (in0 || in1) {
          const uint64_t row = s_base[sl] + i0;
          int64_t t0 = 0, t1 = 0;
          double v0 = 0.0, v1 = 0.0;
          if (row + 1 < a.n_rows) {
            const longlong2 tt = __ldcs(reinterpret_cast<const longlong2*>(a.ts + row));
            const double2 vv = __ldcs(reinterpret_cast<const double2*>(a.val + row));
            t0 = tt.x; t1 = tt.y; v0 = vv.x; v1 = vv.y;
          } else {
            t0 = a.ts[row];
            v0 = a.val[row];
          }
          if (in0) {
            s_ts[(i0 & (kTRing - 1)) * 32 + sl] = TimeDom<true>::conv(t0 + a.offset, a);
            s_val[(i0 & (kTRing - 1)) * 32 + sl] = v0;
            bad |= isnan(v0);
          }
          if (in1) {
            s_ts[((i0 + 1) & (kTRing - 1)) * 32 + sl] = TimeDom<true>::conv(t1 + a.offset, a);
            s_val[((i0 + 1) & (kTRing - 1)) * 32 + sl] = v1;
            bad |= isnan(v1);
          }
        }
        // a NaN row would be dropped by SeriesNormalize and shift every index: not this tier's job
        const uint32_t bad_mask = __ballot_sync(0xffffffffu, bad && a.filter_nan);
        if (bad_mask && (lane >> 2) == g) {
          if ((bad_mask >> ((lane & 3) * 8)) & 0xffu) deferred = true;
        }
      }
      loaded += kTChunk;
      __syncwarp();
      if (deferred) busy = false;
      if (busy) {
        const uint32_t avail = loaded < end_i ? loaded : end_i;
        if (avail > lead) newest = ring.t(avail - 1);
      }
    }

    // ---- hand-offs and series-level rules ---------------------------------------------------------------
    if (have) {
      if (deferred) {
        const uint32_t i = atomicAdd(&a.status->w_count, 1u);
        a.w_list[i] = s;
      } else if (k < T) {
        // unreachable: busy only clears at k == T or on deferral
      }
    }
    __syncwarp();
  }
}

}  // namespace b2p
