// b2p_plan.cpp — host side of GpuPromRangeExec (see b2p_plan.hpp) and its C entry points.
// Pure host C++: everything numeric goes through the C ABI (b2p_range_eval / b2p_group_aggregate).
#include "b2p_plan.hpp"

#include <algorithm>
#include <cctype>
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <unordered_map>

namespace b2p {

namespace {

struct NameId {
  const char* name;
  int id;
};
// UDF display names, src/promql/src/functions/*.rs (`display_name = prom_*`) and planner.rs:2183-2221
const NameId kFns[] = {
    {"prom_rate", B2P_FN_RATE}, {"prom_increase", B2P_FN_INCREASE}, {"prom_delta", B2P_FN_DELTA},
    {"prom_irate", B2P_FN_IRATE}, {"prom_idelta", B2P_FN_IDELTA}, {"prom_resets", B2P_FN_RESETS},
    {"prom_changes", B2P_FN_CHANGES}, {"prom_count_over_time", B2P_FN_COUNT_OVER_TIME},
    {"prom_sum_over_time", B2P_FN_SUM_OVER_TIME}, {"prom_avg_over_time", B2P_FN_AVG_OVER_TIME},
    {"prom_min_over_time", B2P_FN_MIN_OVER_TIME}, {"prom_max_over_time", B2P_FN_MAX_OVER_TIME},
    {"prom_last_over_time", B2P_FN_LAST_OVER_TIME}, {"prom_present_over_time", B2P_FN_PRESENT_OVER_TIME},
    {"prom_absent_over_time", B2P_FN_ABSENT_OVER_TIME}, {"prom_stdvar_over_time", B2P_FN_STDVAR_OVER_TIME},
    {"prom_stddev_over_time", B2P_FN_STDDEV_OVER_TIME}, {"prom_deriv", B2P_FN_DERIV},
    {"prom_predict_linear", B2P_FN_PREDICT_LINEAR}, {"prom_quantile_over_time", B2P_FN_QUANTILE_OVER_TIME},
    {"prom_double_exponential_smoothing", B2P_FN_HOLT_WINTERS}, {"prom_holt_winters", B2P_FN_HOLT_WINTERS},
};
const NameId kAggs[] = {{"sum", B2P_AGG_SUM},       {"avg", B2P_AGG_AVG},       {"count", B2P_AGG_COUNT},
                        {"min", B2P_AGG_MIN},       {"max", B2P_AGG_MAX},       {"stddev", B2P_AGG_STDDEV},
                        {"stdvar", B2P_AGG_STDVAR}};

bool starts_with(const char* s, const char* p) { return std::strncmp(s, p, std::strlen(p)) == 0; }

bool bit_set(const uint8_t* bits, int64_t i) { return bits == nullptr || ((bits[i >> 3] >> (i & 7)) & 1); }

// ---- export helpers: an ArrowArray whose buffers live in a heap object ---------------------------------
struct OwnedColumn {
  std::vector<int64_t> i64;
  std::vector<double> f64;
  std::vector<int32_t> offsets;
  std::string chars;
  const void* buffers[3] = {nullptr, nullptr, nullptr};
};
struct OwnedBatch {
  std::vector<std::unique_ptr<OwnedColumn>> cols;
  std::vector<ArrowArray> child_arrays;
  std::vector<ArrowArray*> child_ptrs;
  const void* buffers[1] = {nullptr};
};
struct OwnedSchema {
  std::vector<std::string> names, formats;
  std::vector<ArrowSchema> children;
  std::vector<ArrowSchema*> child_ptrs;
};

void release_child_array(ArrowArray* a) { a->release = nullptr; }
void release_batch(ArrowArray* a) {
  if (!a || !a->release) return;
  delete static_cast<OwnedBatch*>(a->private_data);
  a->release = nullptr;
}
void release_child_schema(ArrowSchema* s) { s->release = nullptr; }
void release_schema(ArrowSchema* s) {
  if (!s || !s->release) return;
  delete static_cast<OwnedSchema*>(s->private_data);
  s->release = nullptr;
}

}  // namespace

int function_id_from_name(const std::string& n) {
  for (const auto& f : kFns)
    if (n == f.name) return f.id;
  return -1;
}
// str::parse::<f64>() of Rust's std for a label value: optional sign, decimal digits with optional fraction and
// exponent, or inf / infinity / nan in any case — nothing else (no surrounding white space, no hex floats, no locale
// decimal comma: strtod would accept those).  Anything that does not parse is NaN (histogram_fold.rs:791-796).
double parse_f64_like_rust(const std::string& v) {
  const double nan = std::nan("");
  if (v.empty()) return nan;
  size_t i = 0;
  bool neg = false;
  if (v[0] == '+' || v[0] == '-') {
    neg = v[0] == '-';
    i = 1;
  }
  if (i >= v.size()) return nan;
  auto ieq = [&](const char* w) {
    size_t n = std::strlen(w);
    if (v.size() - i != n) return false;
    for (size_t k = 0; k < n; ++k)
      if (std::tolower((unsigned char)v[i + k]) != w[k]) return false;
    return true;
  };
  if (ieq("inf") || ieq("infinity")) return neg ? -HUGE_VAL : HUGE_VAL;
  if (ieq("nan")) return nan;
  bool digits = false, dot = false, exp = false;
  for (size_t k = i; k < v.size(); ++k) {
    const char ch = v[k];
    if (ch >= '0' && ch <= '9') {
      digits = true;
    } else if (ch == '.' && !dot && !exp) {
      dot = true;
    } else if ((ch == 'e' || ch == 'E') && digits && !exp) {
      exp = true;
      if (k + 1 < v.size() && (v[k + 1] == '+' || v[k + 1] == '-')) ++k;
      if (k + 1 >= v.size()) return nan;  // exponent without digits
      digits = true;
      for (size_t q = k + 1; q < v.size(); ++q)
        if (v[q] < '0' || v[q] > '9') return nan;
      break;
    } else {
      return nan;
    }
  }
  if (!digits) return nan;
  double out = 0.0;
  const auto r = std::from_chars(v.data() + i, v.data() + v.size(), out, std::chars_format::general);
  if (r.ec != std::errc() || r.ptr != v.data() + v.size()) {
    if (r.ec == std::errc::result_out_of_range) out = HUGE_VAL;  // Rust saturates to inf (and to 0 on underflow)
    else return nan;
  }
  return neg ? -out : out;
}

int aggregate_id_from_name(const std::string& n) {
  for (const auto& f : kAggs)
    if (n == f.name) return f.id;
  return -1;
}

// ---- RecordBatch -------------------------------------------------------------------------------------
RecordBatch::RecordBatch(ArrowArray* array, ArrowSchema* schema) {
  if (!array || !schema || !array->release || !schema->release)
    throw PlanError(ErrorKind::Internal, "RecordBatch: released or NULL Arrow C structs");
  array_ = *array;
  schema_ = *schema;
  array->release = nullptr;  // moved
  schema->release = nullptr;
  if (!schema_.format || std::strcmp(schema_.format, "+s") != 0 || array_.n_children != schema_.n_children) {
    array_.release(&array_);
    schema_.release(&schema_);
    throw PlanError(ErrorKind::Execution, "RecordBatch: expected a struct array (format \"+s\")");
  }
}
RecordBatch::~RecordBatch() {
  if (array_.release) array_.release(&array_);
  if (schema_.release) schema_.release(&schema_);
}
int RecordBatch::find(const std::string& name) const {
  for (int64_t i = 0; i < schema_.n_children; ++i)
    if (schema_.children[i]->name && name == schema_.children[i]->name) return (int)i;
  return -1;
}

// ---- PromRangePlan -------------------------------------------------------------------------------------
PromRangePlan::PromRangePlan(b2p_ctx* ctx, PromRangePlanArgs args) : ctx_(ctx), args_(std::move(args)) {
  if (!ctx_) throw PlanError(ErrorKind::Internal, "GpuPromRangeExec: NULL context");
  fn_id_ = args_.function.empty() ? -1 : function_id_from_name(args_.function);
  if (fn_id_ < 0 && !args_.function.empty())
    throw PlanError(ErrorKind::Plan, "GpuPromRangeExec: unknown range function " + args_.function);
  if (args_.histogram) {
    if (std::find(args_.tag_columns.begin(), args_.tag_columns.end(), args_.le_column) == args_.tag_columns.end())
      throw PlanError(ErrorKind::Plan, "HistogramFold: le column " + args_.le_column + " is not a tag column");
    if (!args_.aggregate.empty())
      throw PlanError(ErrorKind::Plan, "HistogramFold over an aggregate is not supported by this node");
  }
  agg_id_ = -1;
  if (!args_.aggregate.empty()) {
    agg_id_ = aggregate_id_from_name(args_.aggregate);
    if (agg_id_ < 0) throw PlanError(ErrorKind::Plan, "GpuPromRangeExec: unsupported aggregator " + args_.aggregate);
    for (const auto& b : args_.by_columns)
      if (std::find(args_.tag_columns.begin(), args_.tag_columns.end(), b) == args_.tag_columns.end())
        throw PlanError(ErrorKind::Plan, "GpuPromRangeExec: by-column " + b + " is not a tag column");
  }
  if (args_.interval <= 0) throw PlanError(ErrorKind::Plan, "GpuPromRangeExec: interval must be positive");
  if (args_.time_index.empty() || args_.field_column.empty())
    throw PlanError(ErrorKind::Plan, "GpuPromRangeExec: time index and field column are required");
  tags_.utf8.resize(args_.tag_columns.size());
}

void PromRangePlan::set_histogram(const std::string& le_column, double quantile) {
  if (std::find(args_.tag_columns.begin(), args_.tag_columns.end(), le_column) == args_.tag_columns.end())
    throw PlanError(ErrorKind::Plan, "HistogramFold: le column " + le_column + " is not a tag column");
  if (!args_.aggregate.empty())
    throw PlanError(ErrorKind::Plan, "HistogramFold over an aggregate is not supported by this node");
  args_.histogram = true;
  args_.le_column = le_column;
  args_.quantile = quantile;
}

void PromRangePlan::push(std::unique_ptr<RecordBatch> batch) {
  const RecordBatch& b = *batch;
  const int64_t n = b.num_rows();
  if (n == 0) return;  // an empty batch is skipped (never parks the stream)
  const int ti = b.find(args_.time_index);
  if (ti < 0) throw PlanError(ErrorKind::Plan, "No field named " + args_.time_index);  // field_not_found
  const int fi = b.find(args_.field_column);
  if (fi < 0) throw PlanError(ErrorKind::Plan, "No field named " + args_.field_column);
  const char* tfmt = b.field(ti).format;
  if (!(starts_with(tfmt, "tsm:") || std::strcmp(tfmt, "l") == 0))
    throw PlanError(ErrorKind::Execution, "Time index Column downcast to TimestampMillisecondArray failed");
  if (std::strcmp(b.field(fi).format, "g") != 0)
    throw PlanError(ErrorKind::Execution, "field column " + args_.field_column + " is not Float64");
  const ArrowArray& ta = b.column(ti);
  const ArrowArray& fa = b.column(fi);
  const int64_t* tsv = static_cast<const int64_t*>(ta.buffers[1]) + ta.offset + b.offset();
  const double* fv = static_cast<const double*>(fa.buffers[1]) + fa.offset + b.offset();
  const uint8_t* fvalid = fa.null_count != 0 ? static_cast<const uint8_t*>(fa.buffers[0]) : nullptr;

  // tag columns: Utf8 (int32 offsets) tuple, or a single UInt64 id
  struct TagCol {
    const int32_t* off;
    const char* data;
    const uint8_t* valid;
    int64_t base;
    const uint64_t* ids;
  };
  std::vector<TagCol> tcols;
  for (size_t t = 0; t < args_.tag_columns.size(); ++t) {
    const int ci = b.find(args_.tag_columns[t]);
    if (ci < 0) throw PlanError(ErrorKind::Plan, "No field named " + args_.tag_columns[t]);
    const ArrowArray& ca = b.column(ci);
    const char* fmt = b.field(ci).format;
    TagCol tc{};
    tc.base = ca.offset + b.offset();
    tc.valid = ca.null_count != 0 ? static_cast<const uint8_t*>(ca.buffers[0]) : nullptr;
    if (std::strcmp(fmt, "u") == 0) {
      tc.off = static_cast<const int32_t*>(ca.buffers[1]);
      tc.data = static_cast<const char*>(ca.buffers[2]);
    } else if (std::strcmp(fmt, "L") == 0 && args_.tag_columns.size() == 1) {
      tc.ids = static_cast<const uint64_t*>(ca.buffers[1]);
      key_is_id_ = true;
    } else {
      throw PlanError(ErrorKind::Execution, "tag column " + args_.tag_columns[t] + " must be Utf8 (or one UInt64 id)");
    }
    tcols.push_back(tc);
  }
  auto tag_at = [&](size_t t, int64_t row) -> std::string {
    const TagCol& tc = tcols[t];
    const int64_t r = tc.base + row;
    if (!bit_set(tc.valid, r)) return std::string("\0null", 5);  // NULL label: distinct from every string
    return std::string(tc.data + tc.off[r], (size_t)(tc.off[r + 1] - tc.off[r]));
  };

  // Columns are taken over in bulk (the Arrow values buffers are already the device layout): one memcpy per column
  // and batch, no per-row growth.  SeriesDivide only has to find the rows where a new series starts
  // (find_first_diff_row compares adjacent rows, series_divide.rs:658-667, and the first row of a batch with the last
  // row of the previous one, :636-645); the result is the offsets array the device kernels take — the 4 B/row id
  // column is never built nor shipped.
  const size_t row_base = ts_.size();
  ts_.insert(ts_.end(), tsv, tsv + n);
  val_.insert(val_.end(), fv, fv + n);
  if (fvalid) {  // a NULL field value cannot be inside a window: treat it like the NaN the filter drops
    for (int64_t row = 0; row < n; ++row)
      if (!bit_set(fvalid, fa.offset + b.offset() + row)) val_[row_base + (size_t)row] = std::nan("");
  }
  auto start_series = [&](int64_t row) {
    offsets_.push_back((uint64_t)(row_base + (size_t)row));
    ++num_series_;
  };
  if (key_is_id_) {
    const uint64_t* ids = tcols[0].ids + tcols[0].base;
    int64_t row = 0;
    if (!have_last_ || ids[0] != last_id_) {
      tags_.tsid.push_back(ids[0]);
      start_series(0);
    }
    for (row = 1; row < n; ++row)
      if (ids[row] != ids[row - 1]) {  // (a tight compare loop the compiler vectorises)
        tags_.tsid.push_back(ids[row]);
        start_series(row);
      }
    last_id_ = ids[n - 1];
  } else if (!tcols.empty()) {
    // adjacent-row compare on the raw Utf8 buffers; label strings are only materialised for the first row of a series
    auto same_as_prev = [&](int64_t row) -> bool {  // row >= 1
      for (const TagCol& tc : tcols) {
        const int64_t r = tc.base + row;
        const bool v = bit_set(tc.valid, r), pv = bit_set(tc.valid, r - 1);
        if (v != pv) return false;
        if (!v) continue;
        const int32_t len = tc.off[r + 1] - tc.off[r];
        if (len != tc.off[r] - tc.off[r - 1]) return false;
        if (std::memcmp(tc.data + tc.off[r], tc.data + tc.off[r - 1], (size_t)len) != 0) return false;
      }
      return true;
    };
    auto same_as_last_key = [&]() -> bool {  // first row of this batch against the previous batch's last row
      for (size_t t = 0; t < tcols.size(); ++t)
        if (tag_at(t, 0) != last_key_[t]) return false;
      return true;
    };
    auto open_series = [&](int64_t row) {
      for (size_t t = 0; t < tcols.size(); ++t) tags_.utf8[t].push_back(tag_at(t, row));
      start_series(row);
    };
    if (!have_last_ || !same_as_last_key()) open_series(0);
    for (int64_t row = 1; row < n; ++row)
      if (!same_as_prev(row)) open_series(row);
    last_key_.resize(tcols.size());
    for (size_t t = 0; t < tcols.size(); ++t) last_key_[t] = tag_at(t, n - 1);
  } else if (!have_last_) {
    start_series(0);  // no tag columns: the whole input is one series (series_divide.rs:624-627)
  }
  have_last_ = true;
}

void PromRangePlan::execute(ArrowArray* out, ArrowSchema* out_schema) {
  if (!out || !out_schema) throw PlanError(ErrorKind::Internal, "execute: NULL output structs");
  b2p_range_params p{};
  p.fn_id = fn_id_;
  p.filter_nan = args_.need_filter_out_nan ? 1 : 0;
  p.start = args_.start;
  p.end = args_.end;
  p.interval = args_.interval;
  p.range = args_.range;
  p.offset = args_.offset;
  p.param0 = args_.param0;
  p.param1 = args_.param1;
  const int64_t T = b2p_num_steps(p.start, p.end, p.interval);
  const uint32_t S = (uint32_t)num_series_;
  offsets_.resize((size_t)S);          // (a previous execute() appended the end marker)
  offsets_.push_back((uint64_t)ts_.size());
  const uint32_t Tw = (uint32_t)((T + 31) / 32);
  const bool fold_on_device = args_.histogram && fn_id_ >= 0;  // the dense matrix then never reaches the host
  std::vector<double> dense(fold_on_device ? 0 : (size_t)S * (size_t)T);
  std::vector<uint32_t> valid(fold_on_device ? 0 : (size_t)S * Tw);
  std::vector<int64_t> eval_ts((size_t)T);
  if (S > 0 && T > 0 && !(args_.histogram && fn_id_ >= 0)) {
    int rc;
    if (fn_id_ >= 0) {
      rc = b2p_range_eval(ctx_, &p, ts_.data(), val_.data(), nullptr, offsets_.data(), ts_.size(), S, dense.data(),
                          valid.data(), eval_ts.data());
    } else {  // InstantManipulate
      rc = b2p_instant_select(ctx_, p.start, p.end, p.interval, args_.lookback_delta, p.offset, ts_.data(),
                              val_.data(), nullptr, offsets_.data(), ts_.size(), S, dense.data(), valid.data());
      for (int64_t k = 0; k < T; ++k) eval_ts[(size_t)k] = p.start + k * p.interval;
    }
    if (rc == B2P_E_INVALID || rc == B2P_E_TOO_LARGE) throw PlanError(ErrorKind::Plan, b2p_last_error());
    if (rc == B2P_E_UNSORTED) throw PlanError(ErrorKind::Internal, b2p_last_error());
    if (rc != B2P_OK) throw PlanError(ErrorKind::Execution, b2p_last_error());
  }
  const std::string value_name =
      fn_id_ >= 0 ? args_.function + "(" + args_.time_index + "_range," + args_.field_column + ")" : args_.field_column;

  auto ob = std::make_unique<OwnedBatch>();
  auto os = std::make_unique<OwnedSchema>();
  auto add_col = [&](const std::string& name, const std::string& fmt) -> OwnedColumn* {
    ob->cols.push_back(std::make_unique<OwnedColumn>());
    os->names.push_back(name);
    os->formats.push_back(fmt);
    return ob->cols.back().get();
  };
  int64_t n_out = 0;

  if (args_.histogram) {
    // HistogramFold (histogram_fold.rs:754-820): group the series by their tags without `le`, order each group's
    // buckets by le ascending (parsed as f64, "+Inf" last), one output row per (group, eval ts)
    if (key_is_id_) throw PlanError(ErrorKind::Plan, "HistogramFold needs the le tag column, not a tsid key");
    const size_t le_idx = (size_t)(std::find(args_.tag_columns.begin(), args_.tag_columns.end(), args_.le_column) -
                                   args_.tag_columns.begin());
    // histogram id of every series (hash of the tag tuple without le; ids in first-appearance order), its bound
    std::unordered_map<std::string, uint32_t> hist_ids;
    std::vector<std::vector<std::string>> hist_keys;
    std::vector<uint32_t> hid(S);
    std::vector<double> sle(S);
    std::string flat;
    for (uint32_t s = 0; s < S; ++s) {
      flat.clear();
      for (size_t t = 0; t < args_.tag_columns.size(); ++t)
        if (t != le_idx) {
          flat += tags_.utf8[t][s];
          flat.push_back('\x1f');
        }
      auto it = hist_ids.find(flat);
      if (it == hist_ids.end()) {
        it = hist_ids.emplace(flat, (uint32_t)hist_keys.size()).first;
        std::vector<std::string> key;
        for (size_t t = 0; t < args_.tag_columns.size(); ++t)
          if (t != le_idx) key.push_back(tags_.utf8[t][s]);
        hist_keys.push_back(std::move(key));
      }
      hid[s] = it->second;
      sle[s] = parse_f64_like_rust(tags_.utf8[le_idx][s]);  // le.parse::<f64>().unwrap_or(NaN), histogram_fold.rs:791-796
    }
    const uint32_t H = (uint32_t)hist_keys.size();
    // output order = the reference's: rows sorted by the remaining tags (std::map order of the old implementation)
    std::vector<uint32_t> hist_order(H);
    for (uint32_t h = 0; h < H; ++h) hist_order[h] = h;
    std::sort(hist_order.begin(), hist_order.end(), [&](uint32_t x, uint32_t y) { return hist_keys[x] < hist_keys[y]; });
    std::vector<uint32_t> rank(H);
    for (uint32_t r = 0; r < H; ++r) rank[hist_order[r]] = r;
    // buckets of every histogram in ascending le order, NaN bounds last, ties in scan order (a strict weak ordering)
    std::vector<uint32_t> bucket_series(S);
    for (uint32_t s = 0; s < S; ++s) bucket_series[s] = s;
    std::stable_sort(bucket_series.begin(), bucket_series.end(), [&](uint32_t x, uint32_t y) {
      if (rank[hid[x]] != rank[hid[y]]) return rank[hid[x]] < rank[hid[y]];
      const bool nx = std::isnan(sle[x]), ny = std::isnan(sle[y]);
      if (nx != ny) return ny;
      return !nx && sle[x] < sle[y];
    });
    std::vector<uint32_t> hist_off(H + 1, 0);
    std::vector<double> bucket_le(S);
    for (uint32_t i = 0; i < S; ++i) {
      bucket_le[i] = sle[bucket_series[i]];
      hist_off[rank[hid[bucket_series[i]]] + 1]++;
    }
    for (uint32_t h = 0; h < H; ++h) hist_off[h + 1] += hist_off[h];
    std::vector<double> hq((size_t)H * (size_t)T);
    std::vector<uint32_t> hqv((size_t)H * Tw);
    if (H > 0 && T > 0) {
      if (fn_id_ < 0) throw PlanError(ErrorKind::Plan, "HistogramFold over an instant selector is not supported by this node");
      const int rc = b2p_range_histogram_fold(ctx_, &p, ts_.data(), val_.data(), nullptr, offsets_.data(), ts_.size(), S,
                                              args_.quantile, hist_off.data(), bucket_series.data(), bucket_le.data(), H,
                                              hq.data(), hqv.data());
      if (rc == B2P_E_INVALID || rc == B2P_E_TOO_LARGE) throw PlanError(ErrorKind::Plan, b2p_last_error());
      if (rc == B2P_E_UNSORTED) throw PlanError(ErrorKind::Internal, b2p_last_error());
      if (rc != B2P_OK) throw PlanError(ErrorKind::Execution, b2p_last_error());
      for (int64_t k = 0; k < T; ++k) eval_ts[(size_t)k] = p.start + k * p.interval;
    }
    OwnedColumn* c_ts = add_col(args_.time_index, "tsm:");
    OwnedColumn* c_val = add_col(value_name, "g");
    std::vector<OwnedColumn*> c_tags;
    for (size_t t = 0; t < args_.tag_columns.size(); ++t)
      if (t != le_idx) {
        c_tags.push_back(add_col(args_.tag_columns[t], "u"));
        c_tags.back()->offsets.push_back(0);
      }
    for (uint32_t h = 0; h < H; ++h) {  // rows of histogram hist_order[h] (tag-sorted), one per eval step with a row
      const std::vector<std::string>& key = hist_keys[hist_order[h]];
      for (int64_t k = 0; k < T; ++k) {
        if (!((hqv[(size_t)h * Tw + (size_t)(k >> 5)] >> (k & 31)) & 1u)) continue;
        c_ts->i64.push_back(eval_ts[(size_t)k]);
        c_val->f64.push_back(hq[(size_t)h * (size_t)T + (size_t)k]);
        for (size_t t = 0; t < c_tags.size(); ++t) {
          c_tags[t]->chars += key[t];
          c_tags[t]->offsets.push_back((int32_t)c_tags[t]->chars.size());
        }
        ++n_out;
      }
    }
  } else if (agg_id_ < 0) {
    // rows of Filter(prom_fn IS NOT NULL): {time_index (eval ts), prom_fn(...), tags...}, series-major order
    OwnedColumn* c_ts = add_col(args_.time_index, "tsm:");
    OwnedColumn* c_val = add_col(value_name, "g");
    std::vector<OwnedColumn*> c_tags;
    for (size_t t = 0; t < args_.tag_columns.size(); ++t) {
      c_tags.push_back(add_col(args_.tag_columns[t], key_is_id_ ? "L" : "u"));
      if (!key_is_id_) c_tags.back()->offsets.push_back(0);
    }
    for (uint32_t s = 0; s < S; ++s)
      for (int64_t k = 0; k < T; ++k) {
        if (!((valid[(size_t)s * Tw + (size_t)(k >> 5)] >> (k & 31)) & 1u)) continue;
        c_ts->i64.push_back(eval_ts[(size_t)k]);
        c_val->f64.push_back(dense[(size_t)s * (size_t)T + (size_t)k]);
        for (size_t t = 0; t < c_tags.size(); ++t) {
          if (key_is_id_) {
            c_tags[t]->i64.push_back((int64_t)tags_.tsid[s]);
          } else {
            c_tags[t]->chars += tags_.utf8[t][s];
            c_tags[t]->offsets.push_back((int32_t)c_tags[t]->chars.size());
          }
        }
        ++n_out;
      }
  } else {
    // prom_aggr_expr_to_plan: group keys = by-labels + eval ts; output sorted by (labels asc, ts asc)
    std::vector<size_t> by_idx;
    for (const auto& bname : args_.by_columns)
      by_idx.push_back((size_t)(std::find(args_.tag_columns.begin(), args_.tag_columns.end(), bname) -
                                args_.tag_columns.begin()));
    std::map<std::vector<std::string>, uint32_t> groups;  // ordered => label-sorted output
    std::vector<uint32_t> gid(S);
    for (uint32_t s = 0; s < S; ++s) {
      std::vector<std::string> key;
      for (size_t bi : by_idx) key.push_back(key_is_id_ ? std::to_string(tags_.tsid[s]) : tags_.utf8[bi][s]);
      auto it = groups.find(key);
      if (it == groups.end()) it = groups.emplace(std::move(key), (uint32_t)groups.size()).first;
      gid[s] = it->second;
    }
    const uint32_t G = (uint32_t)groups.size();
    std::vector<double> gval((size_t)G * (size_t)T);
    std::vector<uint32_t> gcnt((size_t)G * (size_t)T);
    if (G > 0 && T > 0) {
      const int rc = b2p_group_aggregate(ctx_, agg_id_, dense.data(), valid.data(), gid.data(), S, G, (uint64_t)T,
                                         gval.data(), gcnt.data());
      if (rc != B2P_OK) throw PlanError(ErrorKind::Execution, b2p_last_error());
    }
    std::vector<OwnedColumn*> c_by;
    for (const auto& bname : args_.by_columns) {
      c_by.push_back(add_col(bname, "u"));
      c_by.back()->offsets.push_back(0);
    }
    OwnedColumn* c_ts = add_col(args_.time_index, "tsm:");
    OwnedColumn* c_val = add_col(args_.aggregate + "(" + (fn_id_ >= 0 ? args_.function : args_.field_column) + ")", "g");
    for (const auto& kv : groups) {  // std::map iterates in key order
      const uint32_t g = kv.second;
      for (int64_t k = 0; k < T; ++k) {
        if (gcnt[(size_t)g * (size_t)T + (size_t)k] == 0) continue;
        for (size_t b = 0; b < c_by.size(); ++b) {
          c_by[b]->chars += kv.first[b];
          c_by[b]->offsets.push_back((int32_t)c_by[b]->chars.size());
        }
        c_ts->i64.push_back(eval_ts[(size_t)k]);
        c_val->f64.push_back(gval[(size_t)g * (size_t)T + (size_t)k]);
        ++n_out;
      }
    }
  }

  // ---- wire up the Arrow C structs ------------------------------------------------------------------
  const size_t nc = ob->cols.size();
  ob->child_arrays.resize(nc);
  ob->child_ptrs.resize(nc);
  os->children.resize(nc);
  os->child_ptrs.resize(nc);
  for (size_t i = 0; i < nc; ++i) {
    OwnedColumn* c = ob->cols[i].get();
    ArrowArray& a = ob->child_arrays[i];
    std::memset(&a, 0, sizeof a);
    a.length = n_out;
    a.null_count = 0;
    a.offset = 0;
    const std::string& fmt = os->formats[i];
    if (fmt == "u") {
      if (c->offsets.empty()) c->offsets.push_back(0);
      c->buffers[0] = nullptr;
      c->buffers[1] = c->offsets.data();
      c->buffers[2] = c->chars.data();
      a.n_buffers = 3;
    } else {
      c->buffers[0] = nullptr;
      c->buffers[1] = (fmt == "g") ? static_cast<const void*>(c->f64.data()) : static_cast<const void*>(c->i64.data());
      a.n_buffers = 2;
    }
    a.buffers = c->buffers;
    a.release = release_child_array;
    ob->child_ptrs[i] = &a;
    ArrowSchema& sc = os->children[i];
    std::memset(&sc, 0, sizeof sc);
    sc.format = os->formats[i].c_str();
    sc.name = os->names[i].c_str();
    sc.flags = 0;
    sc.release = release_child_schema;
    os->child_ptrs[i] = &sc;
  }
  std::memset(out, 0, sizeof *out);
  out->length = n_out;
  out->n_buffers = 1;
  out->buffers = ob->buffers;
  out->n_children = (int64_t)nc;
  out->children = ob->child_ptrs.data();
  out->release = release_batch;
  out->private_data = ob.release();
  std::memset(out_schema, 0, sizeof *out_schema);
  out_schema->format = "+s";
  out_schema->name = "";
  out_schema->n_children = (int64_t)nc;
  out_schema->children = os->child_ptrs.data();
  out_schema->release = release_schema;
  out_schema->private_data = os.release();
}

}  // namespace b2p

// ---- C entry points -----------------------------------------------------------------------------------
struct b2p_plan {
  std::unique_ptr<b2p::PromRangePlan> plan;
};

namespace {
thread_local std::string g_err;
int plan_fail(const b2p::PlanError& e) {
  g_err = e.what();
  switch (e.kind) {
    case b2p::ErrorKind::Plan: return B2P_E_INVALID;
    case b2p::ErrorKind::Internal: return B2P_E_UNSORTED;
    default: return B2P_E_CUDA;
  }
}
}  // namespace

extern "C" {

const char* b2p_plan_last_error(void) { return g_err.c_str(); }

b2p_plan* b2p_plan_range_create(b2p_ctx* ctx, const char* function, const b2p_range_params* p, const char* time_index,
                                const char* field_column, const char* const* tag_columns, int32_t n_tags,
                                const char* aggregate, const char* const* by_columns, int32_t n_by) {
  try {
    if (!function || !p || !time_index || !field_column) throw b2p::PlanError(b2p::ErrorKind::Plan, "NULL argument");
    b2p::PromRangePlanArgs a;
    a.function = function;
    a.start = p->start;
    a.end = p->end;
    a.interval = p->interval;
    a.range = p->range;
    a.offset = p->offset;
    a.need_filter_out_nan = p->filter_nan != 0;
    a.param0 = p->param0;
    a.param1 = p->param1;
    a.time_index = time_index;
    a.field_column = field_column;
    for (int32_t i = 0; i < n_tags; ++i) a.tag_columns.emplace_back(tag_columns[i]);
    if (aggregate && aggregate[0]) a.aggregate = aggregate;
    for (int32_t i = 0; i < n_by; ++i) a.by_columns.emplace_back(by_columns[i]);
    auto* h = new b2p_plan();
    h->plan = std::make_unique<b2p::PromRangePlan>(ctx, std::move(a));
    return h;
  } catch (const b2p::PlanError& e) {
    plan_fail(e);
  } catch (const std::exception& e) {
    g_err = e.what();
  }
  return nullptr;
}

int b2p_plan_set_instant(b2p_plan* plan, int64_t lookback_delta) {
  if (!plan) return B2P_E_INVALID;
  return plan->plan->set_instant(lookback_delta);
}

int b2p_plan_set_histogram_quantile(b2p_plan* plan, const char* le_column, double quantile) {
  if (!plan || !le_column) return B2P_E_INVALID;
  try {
    plan->plan->set_histogram(le_column, quantile);
    return B2P_OK;
  } catch (const b2p::PlanError& e) {
    return plan_fail(e);
  }
}

int b2p_plan_push_batch(b2p_plan* plan, struct ArrowArray* batch, struct ArrowSchema* schema) {
  if (!plan) return B2P_E_INVALID;
  try {
    plan->plan->push(std::make_unique<b2p::RecordBatch>(batch, schema));
    return B2P_OK;
  } catch (const b2p::PlanError& e) {
    return plan_fail(e);
  } catch (const std::exception& e) {
    g_err = e.what();
    return B2P_E_NOMEM;
  }
}

int b2p_plan_execute(b2p_plan* plan, struct ArrowArray* out, struct ArrowSchema* out_schema) {
  if (!plan) return B2P_E_INVALID;
  try {
    plan->plan->execute(out, out_schema);
    return B2P_OK;
  } catch (const b2p::PlanError& e) {
    return plan_fail(e);
  } catch (const std::exception& e) {
    g_err = e.what();
    return B2P_E_NOMEM;
  }
}

int64_t b2p_plan_num_series(b2p_plan* plan) { return plan ? plan->plan->num_series() : -1; }

void b2p_plan_destroy(b2p_plan* plan) { delete plan; }

}  // extern "C"
