// b2p_aggregate.cuh — kernels above the range functions:
//   K3 group_aggregate_kernel  by-label aggregate (DataFusion AggregateExec planned by
//                              prom_aggr_expr_to_plan, src/query/src/promql/planner.rs:334-452)
//   K5 histogram_quantile_kernel  HistogramFold::evaluate_row (histogram_fold.rs:1046-1118)
//   K6 column_reduce_*         per-column sum/count of a wide f64 table (config 5)
#pragma once
#include <cstdint>

#include "b2p_kernels.cuh"

namespace b2p {

// ---------------------------------------------------------------------------------------------
// K3.  Members of group g are series members[goff[g] .. goff[g+1]) in ascending series order
// (stable sort of gid), so every accumulator sees its rows in the same order a single DataFusion
// partition would: plain f64 +=, nulls skipped, group absent (cnt 0) when it receives no row.
// One warp per (group, 32-step tile): each member contributes one coalesced 256-byte segment.
// HBM traffic: reads 8 B + 1 bit per (series, step), writes 12 B per (group, step).
// ---------------------------------------------------------------------------------------------
struct GroupArgs {
  int32_t agg;
  const double* vals;
  const uint32_t* valid;
  const uint32_t* goff;     // [n_groups+1]
  const uint32_t* members;  // [n_member_series]
  uint32_t n_groups;
  uint64_t T;
  uint32_t Tw;
  double* out_val;
  uint32_t* out_cnt;
  int32_t accumulate;  // 1: add into existing out_val/out_cnt (SUM/COUNT partial chaining)
  double* out_mean;    // stddev / stdvar only, may be NULL: when given, out_val receives the raw M2 and out_mean the
                       // mean — the (count, mean, M2) state another rank's partial can be merged with
};

// AGG is a compile-time constant so that the per-member fold is one or two instructions (sum / avg / count) instead
// of a switch inside the inner loop.
template <int AGG>
__global__ void __launch_bounds__(256) group_aggregate_kernel(const GroupArgs a) {
  const int lane = threadIdx.x & 31;
  const uint64_t tiles = (a.T + 31) / 32;
  const uint64_t total = (uint64_t)a.n_groups * tiles;
  const uint32_t tiles32 = (uint32_t)tiles;
  const bool small = total < 0xffffffffull;  // 32-bit task arithmetic (always, in practice)
  for (uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total;
       w += ((uint64_t)gridDim.x * blockDim.x) >> 5) {
    uint32_t g;
    uint64_t tile;
    if (small) {
      g = (uint32_t)w / tiles32;
      tile = (uint32_t)w - g * tiles32;
    } else {
      g = (uint32_t)(w / tiles);
      tile = w - (uint64_t)g * tiles;
    }
    const uint64_t k = tile * 32 + lane;
    const bool in = k < a.T;
    const uint32_t m0 = a.goff[g], m1 = a.goff[g + 1];
    double acc = 0.0, mean = 0.0, m2 = 0.0;
    uint32_t cnt = 0;
    auto fold = [&](double x) {
      if constexpr (AGG == B2P_AGG_SUM || AGG == B2P_AGG_AVG) {
        acc += x;
      } else if constexpr (AGG == B2P_AGG_COUNT) {
      } else if constexpr (AGG == B2P_AGG_MIN) {
        if (cnt == 0 || x < acc || (isnan(acc) && !isnan(x))) acc = x;
      } else if constexpr (AGG == B2P_AGG_MAX) {
        if (cnt == 0 || x > acc || (isnan(acc) && !isnan(x))) acc = x;
      } else {  // Welford, population variance
        const double new_count = (double)cnt + 1.0;
        const double delta1 = x - mean;
        const double new_mean = delta1 / new_count + mean;
        const double delta2 = x - new_mean;
        m2 += delta1 * delta2;
        mean = new_mean;
      }
      ++cnt;
    };
    // Up to 32 members at a time: lane i fetches member i's series id and validity word (two dependent loads for
    // the whole batch instead of two per member), then the 256-byte value segments are requested four members
    // ahead of their use; the fold itself stays in member (= series) order.
    const double* const vals_k = a.vals + k;  // lane's column inside a series' row
    for (uint32_t mb = m0; mb < m1; mb += 32) {
      const uint32_t nm = (m1 - mb < 32u) ? (m1 - mb) : 32u;
      uint32_t s_l = 0, bit_l = 0;
      if ((uint32_t)lane < nm) {
        s_l = a.members[mb + lane];
        bit_l = a.valid[(size_t)s_l * a.Tw + tile];
      }
      for (uint32_t j = 0; j < nm; j += 4) {
        double x[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t src = (j + u < nm) ? j + u : j;  // shuffles stay warp-uniform past the batch end
          const uint32_t s = __shfl_sync(0xffffffffu, s_l, (int)src);
          const uint32_t word = __shfl_sync(0xffffffffu, bit_l, (int)src);
          ok[u] = (j + u < nm) && in && ((word >> lane) & 1u);
          x[u] = 0.0;
          if (ok[u]) x[u] = vals_k[(size_t)s * a.T];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (ok[u]) fold(x[u]);
      }
    }
    if (!in) continue;
    const size_t o = (size_t)g * a.T + k;
    if (a.accumulate) {  // raw partials: SUM-type value and count
      a.out_val[o] += acc;
      a.out_cnt[o] += cnt;
      continue;
    }
    if constexpr (AGG == B2P_AGG_STDVAR || AGG == B2P_AGG_STDDEV) {
      if (a.out_mean) {
        a.out_val[o] = cnt ? m2 : 0.0;
        a.out_mean[o] = cnt ? mean : 0.0;
        a.out_cnt[o] = cnt;
        continue;
      }
    }
    double r = 0.0;
    if (cnt > 0) {
      if constexpr (AGG == B2P_AGG_SUM || AGG == B2P_AGG_MIN || AGG == B2P_AGG_MAX) r = acc;
      else if constexpr (AGG == B2P_AGG_AVG) r = acc / (double)cnt;
      else if constexpr (AGG == B2P_AGG_COUNT) r = (double)cnt;
      else if constexpr (AGG == B2P_AGG_STDVAR) r = m2 / (double)cnt;
      else r = sqrt(m2 / (double)cnt);
    }
    a.out_val[o] = r;
    a.out_cnt[o] = cnt;
  }
}

__global__ void __launch_bounds__(256) group_finalize_kernel(int32_t agg, double* val, const uint32_t* cnt, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t c = cnt[i];
    if (c == 0) { val[i] = 0.0; continue; }
    if (agg == B2P_AGG_AVG) val[i] = val[i] / (double)c;
    else if (agg == B2P_AGG_COUNT) val[i] = (double)c;
    else if (agg == B2P_AGG_STDVAR) val[i] = val[i] / (double)c;        // merged M2 -> population variance
    else if (agg == B2P_AGG_STDDEV) val[i] = sqrt(val[i] / (double)c);
  }
}

// Cross-rank merge helpers of the by-label partials (b2p_allreduce_partials_dev).
// phase 0: groups this rank has no row for become the neutral element of min / max; phase 1 (after the all-reduce,
// cnt = global count): groups absent everywhere read 0.0 again, like a freshly built partial.
__global__ void __launch_bounds__(256) minmax_neutral_kernel(bool is_min, double* val, const uint32_t* cnt, uint64_t n, int phase) {
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    if (cnt[i] == 0) val[i] = phase == 0 ? (is_min ? inf : -inf) : 0.0;
}
// (cnt, mean, M2) states of population variance.  phase 0: wsum = cnt * mean, cnt_r = cnt (kept: cnt becomes global);
// phase 1 (wsum, cnt all-reduced): mean_g = wsum / cnt; M2 += cnt_r * (mean_r - mean_g)^2 — the all-reduce of M2 that
// follows yields the merged M2; mean = mean_g.
__global__ void __launch_bounds__(256) variance_merge_kernel(int phase, double* m2, const uint32_t* cnt, double* mean,
                                                             double* wsum, uint32_t* cnt_r, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    if (phase == 0) {
      cnt_r[i] = cnt[i];
      wsum[i] = (double)cnt[i] * mean[i];
    } else {
      const double mg = cnt[i] ? wsum[i] / (double)cnt[i] : 0.0;
      const double d = mean[i] - mg;
      m2[i] = cnt_r[i] ? m2[i] + (double)cnt_r[i] * d * d : 0.0;
      mean[i] = mg;
    }
  }
}

__global__ void __launch_bounds__(256) iota_kernel(uint32_t* p, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = i;
}

// goff[g] = lower_bound(sorted_gid, g) for g in [0, n_groups]; series with gid >= n_groups fall off the end.
__global__ void __launch_bounds__(256) group_offsets_kernel(const uint32_t* sorted_gid, uint32_t n, uint32_t n_groups,
                                                            uint32_t* goff) {
  for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g <= n_groups; g += gridDim.x * blockDim.x) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (sorted_gid[mid] < g) lo = mid + 1; else hi = mid;
    }
    goff[g] = lo;
  }
}

// ---------------------------------------------------------------------------------------------
// K5.  One warp per (histogram, 32-step tile); lane = step.  Each lane walks the B cumulative
// bucket counters of its (histogram, step) twice — once to get the (monotonised) total, once to
// locate the bucket — reading rates[(h*B+b)*T + k], a coalesced 256-byte segment per bucket.
// A row exists iff all B buckets are present at that step (the reference folds complete groups,
// histogram_fold.rs:772-813).
// ---------------------------------------------------------------------------------------------
struct HistArgs {
  double phi;
  const double* le;  // [B] bucket upper bounds, ascending, last = +Inf
  uint32_t B;
  const double* rates;
  const uint32_t* valid;
  uint32_t n_hist;
  uint64_t T;
  uint32_t Tw;
  double* out;
  uint32_t* out_valid;
};

// The row evaluation of HistogramFold::evaluate_row (histogram_fold.rs:1046-1118) once the monotonised counters
// c[0..B) of the row are known through `get(b)`, its total and the per-query bucket checks.
template <class Get>
__device__ __forceinline__ double histogram_row(const HistArgs& a, bool bucket_sorted, bool last_inf, double total_c, Get get) {
  const double kNaN = __longlong_as_double(0x7ff8000000000000ll);
  const double kInf = __longlong_as_double(0x7ff0000000000000ll);
  if (a.B <= 1) return kNaN;
  if (!last_inf) return kNaN;  // Err("last bucket should be +Inf") -> unwrap_or(NaN), :806
  if (a.phi < 0.0) return -kInf;
  if (a.phi > 1.0) return kInf;
  if (isnan(a.phi)) return kNaN;
  if (!bucket_sorted) return kNaN;
  const double expected_pos = total_c * a.phi;
  // first bucket whose counter is >= expected_pos: the counters are non-decreasing, so bisect
  uint32_t lo = 0, hi = a.B;  // answer in [lo, hi]; hi == B: none (cannot happen: c[B-1] = total >= pos for phi <= 1)
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (get(mid) < expected_pos) lo = mid + 1; else hi = mid;
  }
  const uint32_t fit = lo;
  if (fit >= a.B - 1) return a.le[a.B - 2];
  const double upper_count = get(fit);
  const double upper_bound = a.le[fit];
  double lower_bound = fmin(a.le[0], 0.0), lower_count = 0.0;
  if (fit > 0) {
    lower_bound = a.le[fit - 1];
    lower_count = get(fit - 1);
  }
  if (fabs(upper_count - lower_count) < 1e-10) return kNaN;
  return lower_bound + (upper_bound - lower_bound) / (upper_count - lower_count) * (expected_pos - lower_count);
}

constexpr int kHistWarps = 4;          // warps per CTA of K5
constexpr int kHistSmemBuckets = 64;   // rows with up to this many buckets keep their counters in shared memory

// One pass over HBM: the warp reads the B rate segments of its (histogram, 32-step tile) once (coalesced 256 bytes per
// bucket, eight requests in flight per lane), monotonises them on the way into shared memory ([bucket][lane], conflict
// free), then every lane bisects its own column.  Algorithmic traffic: 8 B x B + B bits read, 8 B + 1 bit written per
// (histogram, step).  Rows with more than kHistSmemBuckets buckets re-read the segments for the bisection (L2 hits).
__global__ void __launch_bounds__(kHistWarps * 32) histogram_quantile_kernel(const HistArgs a) {
  extern __shared__ __align__(16) unsigned char hist_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* col = reinterpret_cast<double*>(hist_smem) + (size_t)warp * kHistSmemBuckets * 32 + lane;  // col[b * 32]
  const uint64_t tiles = (a.T + 31) / 32;
  const uint64_t total = (uint64_t)a.n_hist * tiles;
  const bool in_smem = a.B <= (uint32_t)kHistSmemBuckets;
  // bucket checks are identical for every row (histogram_fold.rs:1047-1073)
  bool bucket_sorted = true;
  for (uint32_t b = 0; b + 1 < a.B; ++b) bucket_sorted &= (a.le[b] <= a.le[b + 1]);
  const bool last_inf = a.B > 0 && !isfinite(a.le[a.B - 1]);
  for (uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total;
       w += ((uint64_t)gridDim.x * blockDim.x) >> 5) {
    const uint32_t h = (uint32_t)(w / tiles);
    const uint64_t tile = w - (uint64_t)h * tiles;
    const uint64_t k = tile * 32 + lane;
    const bool in = k < a.T;
    const size_t s0 = (size_t)h * a.B;
    // a row exists iff all B buckets are present at that step: lanes fetch the validity words of 32 buckets at a time
    uint32_t all = 0xffffffffu;
    for (uint32_t b0 = 0; b0 < a.B; b0 += 32) {
      const uint32_t wv = (b0 + lane < a.B) ? a.valid[(s0 + b0 + lane) * a.Tw + tile] : 0xffffffffu;
      all &= __reduce_and_sync(0xffffffffu, wv);
    }
    const double* seg = a.rates + s0 * a.T + (in ? k : 0);
    double prev = 0.0;
    for (uint32_t b0 = 0; b0 < a.B; b0 += 8) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (b0 + u < a.B && in) ? __ldcs(seg + (size_t)(b0 + u) * a.T) : 0.0;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (b0 + u < a.B) {
          double c = isfinite(v[u]) ? v[u] : prev;  // non-finite -> previous, decreasing -> previous (:1074-1092)
          if (b0 + u > 0 && c < prev) c = prev;
          prev = c;
          if (in_smem) col[(b0 + u) * 32] = c;
        }
      }
    }
    const bool ok = in && ((all >> lane) & 1u);
    double r = 0.0;
    if (ok) {
      if (in_smem) {
        r = histogram_row(a, bucket_sorted, last_inf, prev, [&](uint32_t b) { return col[b * 32]; });
      } else {
        // counters of bucket b on demand: the running maximum needs the buckets before it — walk (rare, wide rows)
        r = histogram_row(a, bucket_sorted, last_inf, prev, [&](uint32_t b) {
          double p = 0.0;
          for (uint32_t q = 0; q <= b; ++q) {
            const double x = seg[(size_t)q * a.T];
            double c = isfinite(x) ? x : p;
            if (q > 0 && c < p) c = p;
            p = c;
          }
          return p;
        });
      }
    }
    if (in) a.out[(size_t)h * a.T + k] = r;
    const uint32_t word = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) a.out_valid[(size_t)h * a.Tw + tile] = word;
  }
}

// ---------------------------------------------------------------------------------------------
// K6.  Deterministic two-stage per-column reduction; NaN rows are skipped like SeriesNormalize's
// filter.  Stage 1: grid (blocks_per_col, n_cols), each block reduces a contiguous slab with 128-bit
// loads; stage 2: one warp per column folds the block partials in order.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) column_reduce_stage1(const double* const* cols, uint64_t n_rows,
                                                            double* part_sum, unsigned long long* part_cnt) {
  const double* col = cols[blockIdx.y];
  const uint64_t per = ((n_rows + gridDim.x - 1) / gridDim.x + 1) & ~1ull;
  const uint64_t r0 = (uint64_t)blockIdx.x * per;
  const uint64_t r1 = min(n_rows, r0 + per);
  double s = 0.0;
  unsigned long long c = 0;
  for (uint64_t r = r0 + 2ull * threadIdx.x; r < r1; r += 2ull * blockDim.x) {
    if (r + 1 < r1) {
      const double2 v = __ldcs(reinterpret_cast<const double2*>(col + r));
      if (!isnan(v.x)) { s += v.x; ++c; }
      if (!isnan(v.y)) { s += v.y; ++c; }
    } else {
      const double v = col[r];
      if (!isnan(v)) { s += v; ++c; }
    }
  }
  __shared__ double ss[8];
  __shared__ unsigned long long sc[8];
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_down_sync(0xffffffffu, s, o);
    c += __shfl_down_sync(0xffffffffu, c, o);
  }
  if ((threadIdx.x & 31) == 0) { ss[threadIdx.x >> 5] = s; sc[threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    unsigned long long tc = 0;
    for (int i = 0; i < 8; ++i) { t += ss[i]; tc += sc[i]; }
    part_sum[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
    part_cnt[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = tc;
  }
}

__global__ void column_reduce_stage2(const double* part_sum, const unsigned long long* part_cnt, uint32_t blocks,
                                     double* out_sum, unsigned long long* out_cnt) {
  const uint32_t c = blockIdx.x;
  if (threadIdx.x != 0) return;
  double t = 0.0;
  unsigned long long tc = 0;
  for (uint32_t i = 0; i < blocks; ++i) { t += part_sum[(size_t)c * blocks + i]; tc += part_cnt[(size_t)c * blocks + i]; }
  out_sum[c] += t;
  out_cnt[c] += tc;
}

}  // namespace b2p
