//! `GpuPromRewrite` — the `PhysicalOptimizerRule` that swaps the PromQL range-query sub-tree for one `GpuPromRangeExec`.
//!
//! Shape matched (bottom-up, exactly what `TQL ANALYZE (0, 10, '5s') rate(test[10s])` prints —
//! tests/cases/standalone/tql-explain-analyze/analyze.result:154-177 — and `tsid_column.result:124-130` for the
//! aggregate on top):
//!
//! ```text
//!   [SortPreservingMergeExec / SortExec(labels, ts)]                                  (planner.rs:443-449)
//!   [AggregateExec(FinalPartitioned) <- RepartitionExec <- AggregateExec(Partial)]    prom_aggr_expr_to_plan
//!   FilterExec: prom_fn(...)@i IS NOT NULL                                            planner.rs:1063
//!   ProjectionExec: expr=[ts, prom_fn(ts_range, val, ts, range_ms) as .., tags..]     planner.rs:1012-1101
//!   PromRangeManipulateExec: req range=[..], interval=[..], eval range=[..]           range_manipulate.rs
//!   PromSeriesNormalizeExec: offset=[..], time index=[..], filter NaN: [..]           normalize.rs
//!   PromSeriesDivideExec: tags=[..]                                                   series_divide.rs
//!   <input: CooperativeExec <- SeriesScan / MergeScanExec>
//! ```
//!
//! Anything that does not match exactly is left alone — the CPU operators keep running for it.  The rule lives in the
//! `promql` crate (src/promql/src/gpu/rule.rs) so that it can read the nodes' fields; the handful of `pub(crate)`
//! getters it needs are listed in `rust-shim/README.md`.
use std::sync::Arc;

use datafusion::common::tree_node::{Transformed, TreeNode};
use datafusion::common::{Result as DataFusionResult, ScalarValue};
use datafusion::config::ConfigOptions;
use datafusion::physical_expr::expressions::{Column, IsNotNullExpr, Literal};
use datafusion::physical_expr::{PhysicalExpr, ScalarFunctionExpr};
use datafusion::physical_optimizer::PhysicalOptimizerRule;
use datafusion::physical_plan::aggregates::{AggregateExec, AggregateMode};
use datafusion::physical_plan::filter::FilterExec;
use datafusion::physical_plan::projection::ProjectionExec;
use datafusion::physical_plan::repartition::RepartitionExec;
use datafusion::physical_plan::ExecutionPlan;

use crate::exec::{GpuPromRangeExec, GpuPromRangeParams};
use crate::ffi::B2pFn;
// In-tree these are `crate::extension_plan::{..}`; named here the way the reference names them.
use promql::extension_plan::{RangeManipulateExec, SeriesDivideExec, SeriesNormalizeExec};

#[derive(Debug)]
pub struct GpuPromRewrite {
    device: i32,
}

impl GpuPromRewrite {
    pub fn new(device: i32) -> Self {
        Self { device }
    }

    /// `FilterExec(prom_fn IS NOT NULL) <- ProjectionExec(prom_fn(..)) <- RangeManipulate <- Normalize <- Divide <- input`
    fn match_range_subtree(&self, plan: &Arc<dyn ExecutionPlan>) -> Option<(GpuPromRangeParams, Arc<dyn ExecutionPlan>)> {
        let filter = plan.as_any().downcast_ref::<FilterExec>()?;
        // predicate: <column i> IS NOT NULL, where column i of the projection is the prom_* call
        let not_null = filter.predicate().as_any().downcast_ref::<IsNotNullExpr>()?;
        let filtered_col = not_null.arg().as_any().downcast_ref::<Column>()?.index();
        let projection = filter.input().as_any().downcast_ref::<ProjectionExec>()?;
        let (udf_expr, _alias) = {
            let e = projection.expr().get(filtered_col)?;
            (e.expr.clone(), e.alias.clone())
        };
        let udf = udf_expr.as_any().downcast_ref::<ScalarFunctionExpr>()?;
        let function = udf.name().to_string();
        B2pFn::from_udf_name(&function)?;
        // every other projected expression must be a plain column (time index, tags): the node emits them unchanged
        for (i, e) in projection.expr().iter().enumerate() {
            if i != filtered_col && e.expr.as_any().downcast_ref::<Column>().is_none() {
                return None;
            }
        }
        let range_exec = projection.input().as_any().downcast_ref::<RangeManipulateExec>()?;
        let normalize = range_exec.input().as_any().downcast_ref::<SeriesNormalizeExec>()?;
        let divide = normalize.input().as_any().downcast_ref::<SeriesDivideExec>()?;
        if range_exec.field_columns().len() != 1 {
            return None; // one value column per series on this path (the reference supports several; they stay on the CPU)
        }
        // UDF arguments (planner.rs:2438-2474): (ts_range, value_range [, ts] [, range_length | scalar params ..])
        let (param0, param1) = scalar_params(&function, udf.args())?;
        let params = GpuPromRangeParams {
            function,
            start: range_exec.start(),
            end: range_exec.end(),
            interval: range_exec.interval(),
            range: range_exec.range(),
            time_index_column: range_exec.time_index_column().to_string(),
            field_column: range_exec.field_columns()[0].clone(),
            offset: normalize.offset(),
            need_filter_out_nan: normalize.need_filter_out_nan(),
            tag_columns: divide.tag_columns().to_vec(),
            param0,
            param1,
            lookback_delta: 0,
            aggregate: None,
            by_columns: vec![],
            histogram: None,
        };
        Some((params, divide.input().clone()))
    }

    /// `AggregateExec(FinalPartitioned) <- RepartitionExec <- AggregateExec(Partial) <- <range sub-tree>` with ONE
    /// aggregate expression among sum / avg / count / min / max / stddev_pop / var_pop and group-by = label columns +
    /// time index (agg_modifier_to_col, planner.rs:1400-1480).
    fn match_aggregate(&self, plan: &Arc<dyn ExecutionPlan>) -> Option<(GpuPromRangeParams, Arc<dyn ExecutionPlan>)> {
        let fin = plan.as_any().downcast_ref::<AggregateExec>()?;
        if !matches!(fin.mode(), AggregateMode::FinalPartitioned | AggregateMode::Final) {
            return None;
        }
        let repart = fin.input().as_any().downcast_ref::<RepartitionExec>()?;
        let partial = repart.input().as_any().downcast_ref::<AggregateExec>()?;
        if !matches!(partial.mode(), AggregateMode::Partial) || partial.aggr_expr().len() != 1 {
            return None;
        }
        let (mut params, input) = self.match_range_subtree(partial.input())?;
        let agg = match partial.aggr_expr()[0].fun().name() {
            "sum" => "sum",
            "avg" => "avg",
            "count" => "count",
            "min" => "min",
            "max" => "max",
            "stddev_pop" => "stddev",
            "var_pop" => "stdvar",
            _ => return None, // quantile / topk / count_values are not all-reduce-able: stay on the CPU
        };
        let mut by = Vec::new();
        for (expr, _name) in partial.group_expr().expr() {
            let col = expr.as_any().downcast_ref::<Column>()?;
            if col.name() != params.time_index_column {
                if !params.tag_columns.iter().any(|t| t == col.name()) {
                    return None;
                }
                by.push(col.name().to_string());
            }
        }
        params.aggregate = Some(agg.to_string());
        params.by_columns = by;
        Some((params, input))
    }
}

/// The scalar UDF arguments the kernels take as (param0, param1); `None` when an argument is not a literal.
fn scalar_params(function: &str, args: &[Arc<dyn PhysicalExpr>]) -> Option<(f64, f64)> {
    let lit = |e: &Arc<dyn PhysicalExpr>| -> Option<f64> {
        match e.as_any().downcast_ref::<Literal>()?.value() {
            ScalarValue::Float64(Some(v)) => Some(*v),
            ScalarValue::Int64(Some(v)) => Some(*v as f64),
            _ => None,
        }
    };
    Some(match function {
        // prom_quantile_over_time(ts_range, value_range, phi); prom_predict_linear(ts_range, value_range, t)
        "prom_quantile_over_time" | "prom_predict_linear" => (lit(args.get(2)?)?, 0.0),
        // prom_holt_winters(ts_range, value_range, sf, tf)
        "prom_holt_winters" | "prom_double_exponential_smoothing" => (lit(args.get(2)?)?, lit(args.get(3)?)?),
        _ => (0.0, 0.0),
    })
}

impl PhysicalOptimizerRule for GpuPromRewrite {
    fn optimize(&self, plan: Arc<dyn ExecutionPlan>, _config: &ConfigOptions) -> DataFusionResult<Arc<dyn ExecutionPlan>> {
        plan.transform_down(|node| {
            // the widest match first: aggregate over the range sub-tree, then the range sub-tree alone
            let matched = self.match_aggregate(&node).or_else(|| self.match_range_subtree(&node));
            match matched {
                Some((params, input)) => {
                    // the replaced node's schema is kept verbatim, so parents (Sort, CoalesceBatches, MergeScan ..) see no change
                    let exec = GpuPromRangeExec::try_new(params, self.device, input, node.schema())?;
                    Ok(Transformed::yes(Arc::new(exec) as Arc<dyn ExecutionPlan>))
                }
                None => Ok(Transformed::no(node)),
            }
        })
        .map(|t| t.data)
    }

    fn name(&self) -> &str {
        "GpuPromRewrite"
    }

    /// The node keeps the schema of what it replaces.
    fn schema_check(&self) -> bool {
        true
    }
}
