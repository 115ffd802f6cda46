//! `GpuPipeTransform` / `GpuGroupByTransform` / `GpuSortTransform`: the processors PipelineBuilder::build adds instead of a
//! Source -> [Filter] -> (Projection | AggregatePartial) [-> Limit] chain (processors/pipeline_builder.rs:26-106).  Each emits
//! exactly what that chain would hand to the next processor — the partial-state Utf8/JSON block
//! (transform_aggregate_partial.rs:61-72) or the filtered + projected (+ limited) rows — so MergeProcessor,
//! AggregateFinalTransform and the final LimitTransform stay unchanged.  UNCOMPILED (see lib.rs); the compiled and tested
//! twin is GpuPipeTransform::execute in fuse_query_b200/csrc/host/pipeline.cc.

use std::collections::HashMap;
use std::ptr;
use std::sync::Arc;

use arrow::array::StringArray;
use async_trait::async_trait;

use fuse_gpu_sys as sys;

use crate::datablocks::DataBlock;
use crate::datasources::Partition;
use crate::datastreams::{DataBlockStream, SendableDataBlockStream};
use crate::datavalues::{DataSchemaRef, DataValue};
use crate::error::{FuseQueryError, FuseQueryResult};
use crate::planners::ExpressionPlan;
use crate::processors::IProcessor;

use super::pipe::{Lowering, Pipe, Source};
use super::table::GpuNumbersTable;
use super::{Column, GpuContext};

pub struct GpuOptions {
    /// numbers_mt: generate in-kernel instead of reading a materialised shard
    pub generated: bool,
    /// let a LIMIT stop the scan (the reference stops pulling blocks, stream_limit.rs:28-62)
    pub limit_early_exit: bool,
    /// reproduce SURVEY F8 (Sum under WHERE with an empty 10 000-row block fails) and the reference's errors under LIMIT
    pub block_quirks: bool,
}

pub struct GpuPipeTransform {
    gpu: Arc<GpuContext>,
    table: Arc<GpuNumbersTable>,
    options: GpuOptions,
    parts: Vec<Partition>,
    predicate: Option<ExpressionPlan>,
    exprs: Vec<ExpressionPlan>,
    is_aggregate: bool,
    schema: DataSchemaRef,
    limit: Option<usize>,
}

/// Leaf states of one select expression in `accumulate_result` order (function_arithmetic.rs:69-75): walks the plan and
/// the node array in the same post-order the lowering pushed them.
fn states_of(e: &ExpressionPlan, next_node: &mut i32, leaf: &HashMap<i32, DataValue>, out: &mut Vec<DataValue>) -> FuseQueryResult<()> {
    match e {
        ExpressionPlan::Alias(_, inner) => {
            states_of(inner, next_node, leaf, out)?;
            *next_node += 1;
        }
        ExpressionPlan::Constant(v) => {
            out.push(v.clone());
            *next_node += 1;
        }
        ExpressionPlan::Field(_) => return Err(FuseQueryError::Internal("Unsupported aggregate operation for function field".to_string())),
        ExpressionPlan::BinaryExpression { left, op, right } => {
            match op.as_str() {
                "+" | "-" | "*" | "/" => {}
                other => return Err(FuseQueryError::Internal(format!("Unsupported aggregate operation for function {}", other))),
            }
            states_of(left, next_node, leaf, out)?;
            states_of(right, next_node, leaf, out)?;
            *next_node += 1;
        }
        ExpressionPlan::Function { args, .. } => {
            skip(&args[0], next_node);
            out.push(leaf.get(next_node).cloned().unwrap_or(DataValue::Null));
            *next_node += 1;
        }
        ExpressionPlan::Wildcard => return Err(FuseQueryError::Internal("Cannot transform wildcard to function".to_string())),
    }
    Ok(())
}

/// nodes the lowering pushed for a subtree that holds no state
fn skip(e: &ExpressionPlan, next_node: &mut i32) {
    match e {
        ExpressionPlan::Alias(_, inner) => skip(inner, next_node),
        ExpressionPlan::BinaryExpression { left, right, .. } => {
            skip(left, next_node);
            skip(right, next_node);
        }
        ExpressionPlan::Function { args, .. } => skip(&args[0], next_node),
        _ => {}
    }
    *next_node += 1;
}

impl GpuPipeTransform {
    #[allow(clippy::too_many_arguments)]
    pub fn try_create(gpu: Arc<GpuContext>, table: Arc<GpuNumbersTable>, options: GpuOptions, parts: Vec<Partition>, predicate: Option<ExpressionPlan>,
                      exprs: Vec<ExpressionPlan>, is_aggregate: bool, schema: DataSchemaRef, limit: Option<usize>) -> FuseQueryResult<Self> {
        // the same construction-time checks as the transforms it replaces (transform_filter.rs:23-29, transform_projection.rs:24-31)
        if let Some(p) = &predicate {
            if p.is_aggregate() {
                return Err(FuseQueryError::Internal(format!("Aggregate function {:?} is found in WHERE in query", p)));
            }
        }
        if !is_aggregate {
            if let Some(e) = exprs.iter().find(|e| e.is_aggregate()) {
                return Err(FuseQueryError::Internal(format!("Unsupported aggregator function: {:?}", e)));
            }
        }
        Ok(GpuPipeTransform { gpu, table, options, parts, predicate, exprs, is_aggregate, schema, limit })
    }

    fn compile(&self, kind: i32) -> FuseQueryResult<(Pipe, Lowering, i32)> {
        let table_schema = crate::datasources::ITable::schema(self.table.as_ref())?;
        let mut lw = Lowering::new(self.options.generated);
        let predicate = match &self.predicate {
            Some(p) => lw.lower(p, &table_schema)?,
            None => -1,
        };
        let first_expr_node = lw.nodes.len() as i32;
        let roots = self.exprs.iter().map(|e| lw.lower(e, &table_schema)).collect::<FuseQueryResult<Vec<_>>>()?;
        let desc = lw.desc(kind, predicate, &roots, &[])?;
        Ok((Pipe::compile(&self.gpu, &desc)?, lw, first_expr_node))
    }

    fn source<'a>(&self, range: super::table::RowRange, col: &'a Column) -> Source<'a> {
        Source { n_rows: range.rows, cols: if self.options.generated { vec![] } else { vec![col] }, generated: self.options.generated, numbers_begin: range.begin }
    }

    fn execute_aggregate(&self) -> FuseQueryResult<DataBlock> {
        let (pipe, _lw, first_expr_node) = self.compile(sys::FQ_PIPE_AGGREGATE)?;
        let has_sum = self.exprs.iter().any(|e| format!("{:?}", e).to_lowercase().contains("sum("));
        let track_blocks = self.options.block_quirks && self.predicate.is_some() && has_sum;
        let runs = self.table.device_runs(&self.parts, track_blocks, ptr::null_mut())?;
        for (k, (range, col)) in runs.iter().enumerate() {
            pipe.launch_aggregate(&self.source(*range, col), k > 0, track_blocks, ptr::null_mut())?;
        }
        let mut leaf = HashMap::new();
        if !runs.is_empty() {
            let (states, _rows) = pipe.fetch_aggregate()?;
            if track_blocks {
                // function_aggregator.rs:88-97: state + arrow_sum(empty block) goes through DataValue::to_array(None)
                let (blocks, empty) = pipe.fetch_block_stats()?;
                if blocks >= 2 && empty > 0 {
                    return Err(FuseQueryError::Internal("DataValue to array cannot be NONE NULL".to_string()));
                }
            }
            leaf = states;
        }
        // one Utf8 row per aggregate expression: serde_json of DataValue::Struct(states) — the wire format of
        // transform_aggregate_partial.rs:61-72, so AggregateFinalTransform parses it unchanged
        let mut next_node = first_expr_node;
        let mut rows = Vec::with_capacity(self.exprs.len());
        for e in &self.exprs {
            let mut states = vec![];
            states_of(e, &mut next_node, &leaf, &mut states)?;
            rows.push(serde_json::to_string(&DataValue::Struct(states)).map_err(|e| FuseQueryError::Internal(e.to_string()))?);
        }
        let col = StringArray::from(rows.iter().map(|s| s.as_str()).collect::<Vec<_>>());
        Ok(DataBlock::create(self.schema.clone(), vec![Arc::new(col)]))
    }

    fn execute_project(&self) -> FuseQueryResult<Vec<DataBlock>> {
        let (pipe, _lw, _) = self.compile(sys::FQ_PIPE_PROJECT)?;
        let exact_errors = self.limit.is_some() && self.options.block_quirks;
        let runs = self.table.device_runs(&self.parts, exact_errors, ptr::null_mut())?;
        let mut out = vec![];
        let mut taken = 0usize;
        for (range, col) in &runs {
            if self.limit == Some(taken) {
                break; // LimitStream ends the pipe (stream_limit.rs:30-31)
            }
            let remaining = self.limit.map(|n| n - taken);
            let capacity = remaining.map_or(range.rows, |n| (n as u64).min(range.rows));
            let (outs, valid) = pipe.alloc_outputs(capacity.max(1))?;
            pipe.launch_project(&self.source(*range, col), &outs, &valid, capacity, remaining, self.options.limit_early_exit, ptr::null_mut())?;
            let fetched = pipe.fetch_project();
            if exact_errors {
                // settle errors over exactly the rows the reference evaluates: everything up to the end of the 10 000-row
                // block FOLLOWING the one that holds the limit-th kept row (LimitStream polls once more, stream_limit.rs:58-62)
                let reached = matches!((&fetched, remaining), (Ok((sel, wr)), Some(n)) if *wr == n as u64 && *sel >= n as u64) || (fetched.is_err() && remaining.is_some());
                let limit_row = if reached { pipe.fetch_limit_row().ok() } else { None };
                match (&fetched, limit_row) {
                    (Err(_), None) => return Err(fetched.unwrap_err()),
                    (_, Some(row)) => {
                        let end = ((row / 10_000 + 2) * 10_000).min(range.rows);
                        let (from, rows) = if fetched.is_err() { (0, end) } else { (row + 1, end - row - 1) };
                        if rows > 0 {
                            let piece = Arc::new(col.slice(from, rows)?);
                            let (o2, v2) = pipe.alloc_outputs(rows)?;
                            let src = Source { n_rows: rows, cols: if self.options.generated { vec![] } else { vec![piece.as_ref()] },
                                               generated: self.options.generated, numbers_begin: range.begin + from };
                            pipe.launch_project(&src, &o2, &v2, rows, None, false, ptr::null_mut())?;
                            pipe.fetch_project()?; // raises iff the reference would
                        }
                        if fetched.is_err() {
                            // the error sat in rows the reference never pulls: rerun the limited launch for its rows
                            pipe.launch_project(&self.source(*range, col), &outs, &valid, capacity, remaining, self.options.limit_early_exit, ptr::null_mut())?;
                            let _ = pipe.fetch_project();
                        }
                    }
                    _ => {}
                }
            }
            let written = match fetched {
                Ok((_, wr)) => wr,
                Err(_) => capacity,
            };
            taken += written as usize;
            let arrays = outs.iter().zip(valid.iter()).map(|(c, v)| c.to_arrow(written, v.as_ref(), ptr::null_mut())).collect::<FuseQueryResult<Vec<_>>>()?;
            out.push(DataBlock::create(self.schema.clone(), arrays));
        }
        Ok(out)
    }
}

#[async_trait]
impl IProcessor for GpuPipeTransform {
    fn name(&self) -> &str {
        "GpuPipeTransform"
    }

    fn connect_to(&mut self, _: Arc<dyn IProcessor>) -> FuseQueryResult<()> {
        Err(FuseQueryError::Internal("Cannot call GpuPipeTransform connect_to".to_string()))
    }

    async fn execute(&self) -> FuseQueryResult<SendableDataBlockStream> {
        let blocks = if self.is_aggregate { vec![self.execute_aggregate()?] } else { self.execute_project()? };
        Ok(Box::pin(DataBlockStream::create(self.schema.clone(), None, blocks)))
    }
}

/// GROUP BY executed as planned (AggregatePlan{group_expr, aggr_expr}, plan_parser.rs:279-308): one processor over all
/// partitions; output = the group fields, then one column per aggregate expression, one row per group.
pub struct GpuGroupByTransform {
    gpu: Arc<GpuContext>,
    table: Arc<GpuNumbersTable>,
    generated: bool,
    parts: Vec<Partition>,
    predicate: Option<ExpressionPlan>,
    group_expr: Vec<ExpressionPlan>,
    aggr_expr: Vec<ExpressionPlan>,
    schema: DataSchemaRef,
}

/// `e` with every aggregate call replaced by a field of the exported leaf block ("__leaf<k>")
fn over_leaves(e: &ExpressionPlan, next_leaf: &mut usize) -> FuseQueryResult<ExpressionPlan> {
    Ok(match e {
        ExpressionPlan::Function { .. } => {
            let name = format!("__leaf{}", *next_leaf);
            *next_leaf += 1;
            ExpressionPlan::Field(name)
        }
        ExpressionPlan::Alias(_, inner) => over_leaves(inner, next_leaf)?,
        ExpressionPlan::Constant(v) => ExpressionPlan::Constant(v.clone()),
        ExpressionPlan::BinaryExpression { left, op, right } => ExpressionPlan::BinaryExpression {
            left: Box::new(over_leaves(left, next_leaf)?),
            op: op.clone(),
            right: Box::new(over_leaves(right, next_leaf)?),
        },
        ExpressionPlan::Field(_) => return Err(FuseQueryError::Internal("Unsupported aggregate operation for function field".to_string())),
        ExpressionPlan::Wildcard => return Err(FuseQueryError::Internal("Cannot transform wildcard to function".to_string())),
    })
}

impl GpuGroupByTransform {
    #[allow(clippy::too_many_arguments)]
    pub fn try_create(gpu: Arc<GpuContext>, table: Arc<GpuNumbersTable>, generated: bool, parts: Vec<Partition>, predicate: Option<ExpressionPlan>,
                      group_expr: Vec<ExpressionPlan>, aggr_expr: Vec<ExpressionPlan>, schema: DataSchemaRef) -> FuseQueryResult<Self> {
        if let Some(k) = group_expr.iter().find(|k| k.is_aggregate()) {
            return Err(FuseQueryError::Internal(format!("Aggregate function {:?} is found in GROUP BY in query", k)));
        }
        Ok(GpuGroupByTransform { gpu, table, generated, parts, predicate, group_expr, aggr_expr, schema })
    }

    fn run(&self) -> FuseQueryResult<DataBlock> {
        let table_schema = crate::datasources::ITable::schema(self.table.as_ref())?;
        let mut lw = Lowering::new(self.generated);
        let predicate = match &self.predicate {
            Some(p) => lw.lower(p, &table_schema)?,
            None => -1,
        };
        let roots = self.aggr_expr.iter().map(|e| lw.lower(e, &table_schema)).collect::<FuseQueryResult<Vec<_>>>()?;
        let keys = self.group_expr.iter().map(|e| lw.lower(e, &table_schema)).collect::<FuseQueryResult<Vec<_>>>()?;
        let pipe = Pipe::compile(&self.gpu, &lw.desc(sys::FQ_PIPE_GROUPBY, predicate, &roots, &keys)?)?;
        let runs = self.table.device_runs(&self.parts, false, ptr::null_mut())?;
        // the table is grown until the groups fit (cardinality is unknown up front)
        let mut hint = 1u64 << 16;
        let groups = loop {
            pipe.groupby_reserve(hint)?;
            for (k, (range, col)) in runs.iter().enumerate() {
                let src = Source { n_rows: range.rows, cols: if self.generated { vec![] } else { vec![col.as_ref()] }, generated: self.generated, numbers_begin: range.begin };
                pipe.launch_groupby(&src, k > 0, ptr::null_mut())?;
            }
            if runs.is_empty() {
                break 0;
            }
            match pipe.fetch_groupby()? {
                Some(n) => break n,
                None => hint *= 8,
            }
        };
        let (kcols, kvalid, lcols, lvalid) = pipe.export_groups(groups, ptr::null_mut())?;
        let mut arrays = vec![];
        for (c, v) in kcols.iter().zip(kvalid.iter()) {
            arrays.push(c.to_arrow(groups, v.as_ref(), ptr::null_mut())?);
        }
        // arithmetic over aggregates, per group, exactly like merge_result re-applies it (function_arithmetic.rs:82-88):
        // a projection pipe over the leaf columns
        if !self.aggr_expr.is_empty() {
            let mut fields = vec![];
            for (k, c) in lcols.iter().enumerate() {
                fields.push(crate::datavalues::DataField::new(&format!("__leaf{}", k), super::dtype_of_tag(c.dtype()), lvalid[k].is_some()));
            }
            let leaf_schema = crate::datavalues::DataSchema::new(fields);
            let mut next_leaf = 0usize;
            let finals = self.aggr_expr.iter().map(|e| over_leaves(e, &mut next_leaf)).collect::<FuseQueryResult<Vec<_>>>()?;
            let mut lw2 = Lowering::new(false);
            let roots2 = finals.iter().map(|e| lw2.lower(e, &leaf_schema)).collect::<FuseQueryResult<Vec<_>>>()?;
            let proj = Pipe::compile(&self.gpu, &lw2.desc(sys::FQ_PIPE_PROJECT, -1, &roots2, &[])?)?;
            let (outs, valid) = proj.alloc_outputs(groups.max(1))?;
            // validity columns of the leaves were attached at export; columns go in pipe order
            let cols: Vec<&Column> = lw2.block_cols.iter().map(|&bi| &lcols[bi]).collect();
            proj.launch_project(&Source { n_rows: groups, cols, generated: false, numbers_begin: 0 }, &outs, &valid, groups, None, false, ptr::null_mut())?;
            proj.fetch_project()?;
            for (c, v) in outs.iter().zip(valid.iter()) {
                arrays.push(c.to_arrow(groups, v.as_ref(), ptr::null_mut())?);
            }
        }
        Ok(DataBlock::create(self.schema.clone(), arrays))
    }
}

#[async_trait]
impl IProcessor for GpuGroupByTransform {
    fn name(&self) -> &str {
        "GpuGroupByTransform"
    }

    fn connect_to(&mut self, _: Arc<dyn IProcessor>) -> FuseQueryResult<()> {
        Err(FuseQueryError::Internal("Cannot call GpuGroupByTransform connect_to".to_string()))
    }

    async fn execute(&self) -> FuseQueryResult<SendableDataBlockStream> {
        let block = self.run()?;
        Ok(Box::pin(DataBlockStream::create(self.schema.clone(), None, vec![block])))
    }
}

/// ORDER BY (no counterpart in the reference: README.md:28 lists sorting as open; sqlparser parses the clause and
/// plan_parser.rs never reads `query.order_by`).  A pipeline breaker: every block of the input is uploaded as one block, the
/// key expressions — over the input's OUTPUT columns — are evaluated by one projection pipe, the row indexes are sorted on the
/// device (ascending unless `descending[j]`, NULLs first, ties in input order) and every column is gathered in that order;
/// with the LIMIT that follows it in the plan only those rows are ordered (radix select) and gathered.
pub struct GpuSortTransform {
    gpu: Arc<GpuContext>,
    keys: Vec<ExpressionPlan>,
    descending: Vec<bool>,
    limit: Option<usize>,
    input: Arc<dyn IProcessor>,
}

impl GpuSortTransform {
    pub fn try_create(gpu: Arc<GpuContext>, keys: Vec<ExpressionPlan>, descending: Vec<bool>, limit: Option<usize>) -> FuseQueryResult<Self> {
        if let Some(k) = keys.iter().find(|k| k.is_aggregate()) {
            return Err(FuseQueryError::Plan(format!("ORDER BY sorts the query's output columns: name the aggregate's column (or its alias) instead of {:?}", k)));
        }
        Ok(GpuSortTransform { gpu, keys, descending, limit, input: Arc::new(crate::processors::EmptyProcessor::create()) })
    }

    fn sort(&self, blocks: Vec<DataBlock>) -> FuseQueryResult<Option<DataBlock>> {
        let blocks: Vec<DataBlock> = blocks.into_iter().filter(|b| b.num_rows() > 0).collect();
        if blocks.is_empty() {
            return Ok(None);
        }
        let schema = blocks[0].schema().clone();
        let n: u64 = blocks.iter().map(|b| b.num_rows() as u64).sum();
        // one device column per field: the blocks' arrays concatenated by arrow, then uploaded with their null bitmaps as they are
        let mut cols = vec![];
        for c in 0..schema.fields().len() {
            let parts: Vec<arrow::array::ArrayRef> = blocks.iter().map(|b| b.column(c).clone()).collect();
            let whole = arrow::compute::concat(&parts).map_err(|e| FuseQueryError::Internal(e.to_string()))?;
            cols.push(Column::from_arrow(&self.gpu, &whole, ptr::null_mut())?);
        }
        // the keys in one projection launch over that block
        let mut lw = Lowering::new(false);
        let roots = self.keys.iter().map(|e| lw.lower(e, &schema)).collect::<FuseQueryResult<Vec<_>>>()?;
        let proj = Pipe::compile(&self.gpu, &lw.desc(sys::FQ_PIPE_PROJECT, -1, &roots, &[])?)?;
        let (key_cols, key_valid) = proj.alloc_outputs(n)?;
        let inputs: Vec<&Column> = lw.block_cols.iter().map(|&bi| &cols[bi]).collect();
        proj.launch_project(&Source { n_rows: n, cols: inputs, generated: false, numbers_begin: 0 }, &key_cols, &key_valid, n, None, false, ptr::null_mut())?;
        proj.fetch_project()?;
        for (k, v) in key_cols.iter().zip(key_valid.iter()) {
            if let Some(v) = v {
                k.set_validity(v)?;     // the sort reads a key's validity from the column itself
            }
        }
        let key_refs: Vec<&Column> = key_cols.iter().collect();
        let (rows, count) = match self.limit {
            Some(l) => Column::sort_indices_limit(&self.gpu, &key_refs, &self.descending, n, l as u64, ptr::null_mut())?,
            None => (Column::sort_indices(&self.gpu, &key_refs, &self.descending, n, ptr::null_mut())?, n),
        };
        if count == 0 {
            return Ok(None);
        }
        let mut arrays = vec![];
        for c in cols.iter() {
            let taken = c.take(&rows, count, c.is_nullable(), ptr::null_mut())?;
            arrays.push(taken.to_arrow(count, taken.validity(), ptr::null_mut())?);
        }
        Ok(Some(DataBlock::create(schema, arrays)))
    }
}

#[async_trait]
impl IProcessor for GpuSortTransform {
    fn name(&self) -> &str {
        "GpuSortTransform"
    }

    fn connect_to(&mut self, input: Arc<dyn IProcessor>) -> FuseQueryResult<()> {
        self.input = input;
        Ok(())
    }

    async fn execute(&self) -> FuseQueryResult<SendableDataBlockStream> {
        use futures::stream::StreamExt;
        let mut stream = self.input.execute().await?;
        let mut blocks = vec![];
        let mut schema: Option<DataSchemaRef> = None;
        while let Some(block) = stream.next().await {
            let block = block?;
            schema.get_or_insert_with(|| block.schema().clone());
            blocks.push(block);
        }
        let schema = schema.unwrap_or_else(|| Arc::new(crate::datavalues::DataSchema::empty()));
        let out = self.sort(blocks)?.into_iter().collect::<Vec<_>>();
        Ok(Box::pin(DataBlockStream::create(schema, None, out)))
    }
}
