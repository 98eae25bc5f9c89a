// b2p_plan.hpp — C++17 host side above the C ABI: the reference's operator interface for this path,
// over the Arrow C Data Interface (what arrow-rs `FFI_ArrowArray` / pyarrow `_export_to_c` produce).
//
// The reference plans    SeriesDivide(tag_columns, time_index)            series_divide.rs:83-110
//                     -> SeriesNormalize(offset, time_index, need_filter_out_nan, tag_columns)  normalize.rs:66-83
//                     -> RangeManipulate(start, end, interval, range, time_index, field_columns) range_manipulate.rs:86-110
//                     -> Projection(prom_xxx(ts_range, field, ts, range))   src/query/src/promql/planner.rs:1012-1101
//                     -> Filter(field IS NOT NULL)                           planner.rs:1063
//                    [-> Aggregate(by-labels + ts, [sum|avg|count|min|max|stddev|stdvar](field)).sort(...)  planner.rs:334-452]
// PromRangePlan is that whole sub-tree as ONE node: same constructor arguments (names and meaning),
// input = RecordBatches sorted by (tag columns, time index) like SeriesDivideExec requires
// (series_divide.rs:410-440), output = the rows the reference's Filter (or Aggregate+Sort) would emit.
// Errors mirror DataFusionError kinds: Plan (bad arguments / missing column, like field_not_found,
// range_manipulate.rs:127-133), Execution (wrong column type, range_manipulate.rs:700-705), Internal.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b200promql.h"

namespace b2p {

enum class ErrorKind { Plan, Execution, Internal };
struct PlanError : std::runtime_error {
  ErrorKind kind;
  PlanError(ErrorKind k, const std::string& m) : std::runtime_error(m), kind(k) {}
};

using Millisecond = int64_t;  // extension_plan.rs:42

// An imported Arrow struct array (= RecordBatch); owns the C structs and releases them.
class RecordBatch {
 public:
  RecordBatch(ArrowArray* array, ArrowSchema* schema);  // moves *array / *schema in
  ~RecordBatch();
  RecordBatch(const RecordBatch&) = delete;
  RecordBatch& operator=(const RecordBatch&) = delete;
  int64_t num_rows() const { return array_.length; }
  int find(const std::string& name) const;  // -1 when absent
  const ArrowArray& column(int i) const { return *array_.children[i]; }
  const ArrowSchema& field(int i) const { return *schema_.children[i]; }
  int64_t offset() const { return array_.offset; }

 private:
  ArrowArray array_;
  ArrowSchema schema_;
};

struct PromRangePlanArgs {
  // prom_* UDF name exactly as the planner writes it ("prom_rate", "prom_avg_over_time", ... planner.rs:2183-2221)
  std::string function;
  // RangeManipulate::new
  Millisecond start = 0, end = 0, interval = 0, range = 0;
  std::string time_index;
  std::string field_column;
  // SeriesNormalize::new
  Millisecond offset = 0;
  bool need_filter_out_nan = true;
  // SeriesDivide::new — Utf8 tag columns, or one UInt64 column (__tsid, TagIdentifier::Id)
  std::vector<std::string> tag_columns;
  // UDF scalar arguments (quantile phi / predict_linear t / smoothing sf, tf)
  double param0 = 0.0, param1 = 0.0;
  // optional prom_aggr_expr_to_plan stage: "", "sum", "avg", "count", "min", "max", "stddev", "stdvar"
  std::string aggregate;
  std::vector<std::string> by_columns;  // must be a subset of tag_columns
  // function == "" selects the instant-vector form instead: InstantManipulate::new(start, end, lookback_delta,
  // interval, time_index, field_column) (instant_manipulate.rs:189-208); `range` is ignored
  Millisecond lookback_delta = 300000;
  // optional HistogramFold::new(le_column, field, time_index, quantile) on top (histogram_fold.rs:104-130):
  // le_column must be one of tag_columns; every histogram must expose the same bucket bounds
  bool histogram = false;
  std::string le_column;
  double quantile = 0.0;
};

class PromRangePlan {
 public:
  PromRangePlan(b2p_ctx* ctx, PromRangePlanArgs args);
  const char* name() const { return "GpuPromRangeExec"; }
  // input stream, in order; batches of one partition (sorted by tags, ts)
  void push(std::unique_ptr<RecordBatch> batch);
  // runs the sub-plan on the device and exports the result batch (caller releases it)
  void execute(ArrowArray* out, ArrowSchema* out_schema);
  int64_t num_series() const { return num_series_; }  // the reference's `num_series` metric (range_manipulate.rs:610-619)
  // switch the node to the instant-vector form (InstantManipulate) / add a HistogramFold on top; before execute()
  int set_instant(Millisecond lookback_delta) {
    args_.function.clear();
    fn_id_ = -1;
    args_.lookback_delta = lookback_delta;
    return 0;
  }
  void set_histogram(const std::string& le_column, double quantile);

 private:
  struct TagStore {
    std::vector<std::vector<std::string>> utf8;  // [tag][series] label values of each series' first row
    std::vector<uint64_t> tsid;                 // when the key is a UInt64 id
  };
  b2p_ctx* ctx_;
  PromRangePlanArgs args_;
  int fn_id_;
  int agg_id_;
  std::vector<int64_t> ts_;
  std::vector<double> val_;
  std::vector<uint64_t> offsets_;  // first row of every series (SeriesDivide's output), end marker added by execute()
  TagStore tags_;
  bool key_is_id_ = false;
  int64_t num_series_ = 0;
  // last row's key, to continue a series across batch boundaries (series_divide.rs:636-645)
  std::vector<std::string> last_key_;
  uint64_t last_id_ = 0;
  bool have_last_ = false;
};

int function_id_from_name(const std::string& prom_name);  // -1 when unknown
int aggregate_id_from_name(const std::string& name);      // -1 when unknown

}  // namespace b2p
