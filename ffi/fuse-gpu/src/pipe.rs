//! ExpressionPlan -> fq_expr_node[] lowering and the `Pipe` handle (one fused kernel per
//! Source -> [Filter] -> (Projection | AggregatePartial | GROUP BY) [-> Limit] chain, pipeline_builder.rs:26-106).
//! UNCOMPILED (see lib.rs).  Mirrors `Lowering` + `compile_pipe` in fuse_query_b200/csrc/host/functions.cc.

use std::collections::HashMap;
use std::os::raw::c_void;
use std::ptr;
use std::sync::Arc;

use fuse_gpu_sys as sys;

use crate::datavalues::{DataSchema, DataValue};
use crate::error::{FuseQueryError, FuseQueryResult};
use crate::planners::ExpressionPlan;

use super::{check, dtype_tag, value_of, Column, GpuContext};

/// Flat node array shared by a pipe's predicate, select expressions and GROUP BY keys.
#[derive(Default)]
pub struct Lowering {
    pub nodes: Vec<sys::fq_expr_node>,
    /// block column index of every pipe input column, in pipe order
    pub block_cols: Vec<usize>,
    pub col_dtypes: Vec<sys::fq_dtype>,
    pub col_nullable: Vec<i32>,
    pub generated: bool,
    /// node index of every Aggregator leaf, keyed by its Debug string position in visit order
    pub agg_nodes: Vec<i32>,
}

fn node(kind: i32, op: i32, left: i32, right: i32) -> sys::fq_expr_node {
    sys::fq_expr_node { kind, op, left, right, column: 0, dtype: sys::FQ_NULL, value: sys::fq_scalar_bits { u: 0 } }
}

impl Lowering {
    /// `generated`: the source is system.numbers_mt computed in-kernel; its only column is pipe column 0.
    pub fn new(generated: bool) -> Self {
        let mut lw = Lowering::default();
        if generated {
            lw.generated = true;
            lw.block_cols.push(0);
            lw.col_dtypes.push(sys::FQ_U64);
            lw.col_nullable.push(0);
        }
        lw
    }

    fn column_of(&mut self, schema: &DataSchema, name: &str) -> FuseQueryResult<i32> {
        let bi = schema.index_of(name)?;
        if let Some(k) = self.block_cols.iter().position(|&c| c == bi) {
            return Ok(k as i32);
        }
        if self.block_cols.len() == sys::FQ_MAX_COLS as usize {
            return Err(FuseQueryError::Internal("Unsupported on the device path: more than 8 input columns in one expression".to_string()));
        }
        let field = schema.field(bi);
        self.block_cols.push(bi);
        self.col_dtypes.push(dtype_tag(field.data_type())?);
        self.col_nullable.push(if field.is_nullable() { 2 } else { 0 });   // 2: validity = arrow's bitmap, read in place
        Ok(self.block_cols.len() as i32 - 1)
    }

    /// Post-order walk of the plan; operator names are those of ScalarFunctionFactory::get (function_factory.rs:17-39).
    pub fn lower(&mut self, e: &ExpressionPlan, schema: &DataSchema) -> FuseQueryResult<i32> {
        let n = match e {
            ExpressionPlan::Field(name) => {
                let mut n = node(sys::FQ_EXPR_FIELD, 0, -1, -1);
                n.column = self.column_of(schema, name)?;
                n
            }
            ExpressionPlan::Constant(v) => {
                let mut n = node(sys::FQ_EXPR_CONSTANT, 0, -1, -1);
                let (dtype, bits) = constant_bits(v)?;
                n.dtype = dtype;
                n.value = bits;
                n
            }
            ExpressionPlan::Alias(_, inner) => {
                let l = self.lower(inner, schema)?;
                node(sys::FQ_EXPR_ALIAS, 0, l, -1)
            }
            ExpressionPlan::BinaryExpression { left, op, right } => {
                let l = self.lower(left, schema)?;
                let r = self.lower(right, schema)?;
                let (kind, code) = binary_op(op)?;
                node(kind, code, l, r)
            }
            ExpressionPlan::Function { op, args } => {
                let code = match op.to_lowercase().as_str() {
                    "min" => sys::FQ_AGG_MIN,
                    "max" => sys::FQ_AGG_MAX,
                    "sum" => sys::FQ_AGG_SUM,
                    "count" => sys::FQ_AGG_COUNT,
                    _ => return Err(FuseQueryError::Internal(format!("Unsupported Function: {}", op))),
                };
                if args.is_empty() {
                    return Err(FuseQueryError::Internal("index out of bounds: the len is 0 but the index is 0".to_string()));
                }
                let l = self.lower(&args[0], schema)?;
                node(sys::FQ_EXPR_AGGREGATOR, code, l, -1)
            }
            ExpressionPlan::Wildcard => return Err(FuseQueryError::Internal("Cannot transform wildcard to function".to_string())),
        };
        self.nodes.push(n);
        let idx = self.nodes.len() as i32 - 1;
        if n.kind == sys::FQ_EXPR_AGGREGATOR {
            self.agg_nodes.push(idx);
        }
        Ok(idx)
    }

    pub fn desc(&self, kind: i32, predicate: i32, roots: &[i32], keys: &[i32]) -> FuseQueryResult<sys::fq_pipe_desc> {
        if roots.len() > sys::FQ_MAX_EXPRS as usize || keys.len() > sys::FQ_MAX_KEYS as usize {
            return Err(FuseQueryError::Internal("Unsupported on the device path: more than 8 select expressions (or 4 keys) in one pipe".to_string()));
        }
        let mut d = sys::fq_pipe_desc {
            n_cols: self.col_dtypes.len() as i32,
            col_dtypes: [sys::FQ_NULL; sys::FQ_MAX_COLS as usize],
            col_nullable: [0; sys::FQ_MAX_COLS as usize],
            generated: self.generated as i32,
            nodes: self.nodes.as_ptr(),
            n_nodes: self.nodes.len() as i32,
            predicate,
            kind,
            n_exprs: roots.len() as i32,
            exprs: [0; sys::FQ_MAX_EXPRS as usize],
            n_keys: keys.len() as i32,
            keys: [0; sys::FQ_MAX_KEYS as usize],
        };
        for (k, t) in self.col_dtypes.iter().enumerate() {
            d.col_dtypes[k] = *t;
            d.col_nullable[k] = self.col_nullable[k];
        }
        d.exprs[..roots.len()].copy_from_slice(roots);
        d.keys[..keys.len()].copy_from_slice(keys);
        Ok(d)
    }
}

fn binary_op(op: &str) -> FuseQueryResult<(i32, i32)> {
    Ok(match op.to_lowercase().as_str() {
        "+" => (sys::FQ_EXPR_ARITHMETIC, sys::FQ_AR_ADD),
        "-" => (sys::FQ_EXPR_ARITHMETIC, sys::FQ_AR_SUB),
        "*" => (sys::FQ_EXPR_ARITHMETIC, sys::FQ_AR_MUL),
        "/" => (sys::FQ_EXPR_ARITHMETIC, sys::FQ_AR_DIV),
        "=" => (sys::FQ_EXPR_COMPARISON, sys::FQ_CMP_EQ),
        "<" => (sys::FQ_EXPR_COMPARISON, sys::FQ_CMP_LT),
        "<=" => (sys::FQ_EXPR_COMPARISON, sys::FQ_CMP_LTEQ),
        ">" => (sys::FQ_EXPR_COMPARISON, sys::FQ_CMP_GT),
        ">=" => (sys::FQ_EXPR_COMPARISON, sys::FQ_CMP_GTEQ),
        "and" => (sys::FQ_EXPR_LOGIC, sys::FQ_LG_AND),
        "or" => (sys::FQ_EXPR_LOGIC, sys::FQ_LG_OR),
        other => return Err(FuseQueryError::Internal(format!("Unsupported Function: {}", other))),
    })
}

fn constant_bits(v: &DataValue) -> FuseQueryResult<(sys::fq_dtype, sys::fq_scalar_bits)> {
    use sys::fq_scalar_bits as B;
    // DataValue::to_array refuses Type(None) (data_value.rs:104-109)
    let none = || FuseQueryError::Internal(format!("DataValue to array cannot be NONE {:?}", v));
    Ok(match v {
        DataValue::Boolean(x) => (sys::FQ_BOOL, B { i: x.ok_or_else(none)? as i64 }),
        DataValue::Int8(x) => (sys::FQ_I8, B { i: x.ok_or_else(none)? as i64 }),
        DataValue::Int16(x) => (sys::FQ_I16, B { i: x.ok_or_else(none)? as i64 }),
        DataValue::Int32(x) => (sys::FQ_I32, B { i: x.ok_or_else(none)? as i64 }),
        DataValue::Int64(x) => (sys::FQ_I64, B { i: x.ok_or_else(none)? }),
        DataValue::UInt8(x) => (sys::FQ_U8, B { u: x.ok_or_else(none)? as u64 }),
        DataValue::UInt16(x) => (sys::FQ_U16, B { u: x.ok_or_else(none)? as u64 }),
        DataValue::UInt32(x) => (sys::FQ_U32, B { u: x.ok_or_else(none)? as u64 }),
        DataValue::UInt64(x) => (sys::FQ_U64, B { u: x.ok_or_else(none)? }),
        DataValue::Float32(x) => (sys::FQ_F32, B { f: x.ok_or_else(none)? as f64 }),
        DataValue::Float64(x) => (sys::FQ_F64, B { f: x.ok_or_else(none)? }),
        other => return Err(FuseQueryError::Internal(format!("Unsupported on the device path: constant {:?}", other))),
    })
}

/// What a launch reads: materialised columns in pipe order, or the in-kernel numbers generator.
pub struct Source<'a> {
    pub n_rows: u64,
    pub cols: Vec<&'a Column>,
    pub generated: bool,
    pub numbers_begin: u64,
}

/// RAII handle of a compiled pipe.  Driven by one thread at a time (it owns running state), like the reference clones its
/// Function per pipe (pipeline_builder.rs:50-65).
pub struct Pipe {
    ctx: Arc<GpuContext>,
    raw: *mut sys::fq_pipe,
    pub kind: i32,
    pub n_exprs: usize,
    pub n_keys: usize,
}

unsafe impl Send for Pipe {}

impl Pipe {
    pub fn compile(ctx: &Arc<GpuContext>, desc: &sys::fq_pipe_desc) -> FuseQueryResult<Self> {
        let mut raw = ptr::null_mut();
        check(ctx.raw, unsafe { sys::fq_pipe_compile(ctx.raw, desc, &mut raw) })?;
        Ok(Pipe { ctx: ctx.clone(), raw, kind: desc.kind, n_exprs: desc.n_exprs as usize, n_keys: desc.n_keys as usize })
    }

    pub(crate) fn raw(&self) -> *mut sys::fq_pipe {
        self.raw
    }

    fn with_source<R>(&self, src: &Source, f: impl FnOnce(&sys::fq_source) -> R) -> R {
        let ptrs: Vec<*const sys::fq_column> = src.cols.iter().map(|c| c.raw as *const sys::fq_column).collect();
        let s = sys::fq_source {
            n_rows: src.n_rows,
            n_cols: ptrs.len() as i32,
            generated: src.generated as i32,
            cols: if ptrs.is_empty() { ptr::null() } else { ptrs.as_ptr() },
            numbers_begin: src.numbers_begin,
        };
        f(&s)
    }

    pub fn expr_dtype(&self, i: usize) -> FuseQueryResult<sys::fq_dtype> {
        let mut t = sys::FQ_NULL;
        check(self.ctx.raw, unsafe { sys::fq_pipe_expr_dtype(self.ctx.raw, self.raw, i as i32, &mut t) })?;
        Ok(t)
    }

    pub fn expr_nullable(&self, i: usize) -> FuseQueryResult<bool> {
        let mut n = 0;
        check(self.ctx.raw, unsafe { sys::fq_pipe_expr_nullable(self.ctx.raw, self.raw, i as i32, &mut n) })?;
        Ok(n != 0)
    }

    // ---- AggregatePartial: Function::accumulate over a whole shard (function_aggregator.rs:57-100) ----
    pub fn launch_aggregate(&self, src: &Source, accumulate: bool, block_stats: bool, stream: *mut c_void) -> FuseQueryResult<()> {
        let flags = (if accumulate { sys::FQ_RUN_ACCUMULATE } else { 0 } | if block_stats { sys::FQ_RUN_BLOCK_STATS } else { 0 }) as u32;
        self.with_source(src, |s| check(self.ctx.raw, unsafe { sys::fq_pipe_launch_aggregate(self.ctx.raw, self.raw, s, flags, stream) }))
    }

    /// node index of every Aggregator leaf, in the order the states come back
    pub fn aggregator_nodes(&self) -> FuseQueryResult<Vec<i32>> {
        let mut nodes = vec![0i32; 64];
        let mut n = 0;
        check(self.ctx.raw, unsafe { sys::fq_pipe_aggregator_nodes(self.ctx.raw, self.raw, nodes.as_mut_ptr(), 64, &mut n) })?;
        nodes.truncate(n as usize);
        Ok(nodes)
    }

    fn fetch_states(&self, merged: bool) -> FuseQueryResult<(HashMap<i32, DataValue>, u64)> {
        let zero = sys::fq_value { dtype: sys::FQ_NULL, some: 0, v: sys::fq_scalar_bits { u: 0 } };
        let mut vals = vec![zero; 64];
        let (mut n, mut rows) = (0i32, 0u64);
        let st = unsafe {
            if merged {
                sys::fq_pipe_fetch_merged(self.ctx.raw, self.raw, vals.as_mut_ptr(), 64, &mut n, &mut rows)
            } else {
                sys::fq_pipe_fetch_aggregate(self.ctx.raw, self.raw, vals.as_mut_ptr(), 64, &mut n, &mut rows)
            }
        };
        check(self.ctx.raw, st)?;
        let nodes = self.aggregator_nodes()?;
        Ok((nodes.iter().zip(vals.iter()).map(|(k, v)| (*k, value_of(v))).collect(), rows))
    }

    /// Aggregator node index -> the state `accumulate_result` would hold, and the post-filter row count.
    pub fn fetch_aggregate(&self) -> FuseQueryResult<(HashMap<i32, DataValue>, u64)> {
        self.fetch_states(false)
    }

    /// The same for the state merged over every rank of the pipe's group by the last launch.
    pub fn fetch_merged(&self) -> FuseQueryResult<(HashMap<i32, DataValue>, u64)> {
        self.fetch_states(true)
    }

    /// (reference 10 000-row blocks scanned, of which the predicate emptied) — SURVEY F8
    pub fn fetch_block_stats(&self) -> FuseQueryResult<(u64, u64)> {
        let (mut blocks, mut empty) = (0u64, 0u64);
        check(self.ctx.raw, unsafe { sys::fq_pipe_fetch_block_stats(self.ctx.raw, self.raw, &mut blocks, &mut empty) })?;
        Ok((blocks, empty))
    }

    // ---- Filter + Projection (+ Limit) ----
    /// -> (output columns, validity columns of the nullable ones)
    pub fn alloc_outputs(&self, capacity: u64) -> FuseQueryResult<(Vec<Column>, Vec<Option<Column>>)> {
        let mut outs = Vec::with_capacity(self.n_exprs);
        let mut valid = Vec::with_capacity(self.n_exprs);
        for i in 0..self.n_exprs {
            outs.push(Column::alloc(&self.ctx, self.expr_dtype(i)?, capacity)?);
            valid.push(if self.expr_nullable(i)? { Some(Column::alloc(&self.ctx, sys::FQ_BOOL, capacity)?) } else { None });
        }
        Ok((outs, valid))
    }

    pub fn launch_project(&self, src: &Source, outs: &[Column], valid: &[Option<Column>], capacity: u64, limit: Option<usize>, early_exit: bool,
                          stream: *mut c_void) -> FuseQueryResult<()> {
        let o: Vec<*mut sys::fq_column> = outs.iter().map(|c| c.raw).collect();
        let v: Vec<*mut sys::fq_column> = valid.iter().map(|c| c.as_ref().map_or(ptr::null_mut(), |c| c.raw)).collect();
        let flags = if early_exit { sys::FQ_RUN_LIMIT_EARLY_EXIT as u32 } else { 0 };
        let lim = limit.map_or(-1i64, |n| n as i64);
        self.with_source(src, |s| {
            check(self.ctx.raw, unsafe { sys::fq_pipe_launch_project(self.ctx.raw, self.raw, s, o.as_ptr(), v.as_ptr(), capacity, lim, flags, stream) })
        })
    }

    /// -> (rows selected, rows written); an evaluation error (zero divisor) surfaces here
    pub fn fetch_project(&self) -> FuseQueryResult<(u64, u64)> {
        let (mut sel, mut wr) = (0u64, 0u64);
        check(self.ctx.raw, unsafe { sys::fq_pipe_fetch_project(self.ctx.raw, self.raw, &mut sel, &mut wr) })?;
        Ok((sel, wr))
    }

    /// source row that produced the last output row of a launch that filled its capacity (completes the LIMIT)
    pub fn fetch_limit_row(&self) -> FuseQueryResult<u64> {
        let mut row = 0u64;
        check(self.ctx.raw, unsafe { sys::fq_pipe_fetch_limit_row(self.ctx.raw, self.raw, &mut row) })?;
        Ok(row)
    }

    // ---- GROUP BY ----
    pub fn groupby_reserve(&self, groups: u64) -> FuseQueryResult<()> {
        check(self.ctx.raw, unsafe { sys::fq_pipe_groupby_reserve(self.ctx.raw, self.raw, groups) })
    }

    pub fn launch_groupby(&self, src: &Source, accumulate: bool, stream: *mut c_void) -> FuseQueryResult<()> {
        let flags = if accumulate { sys::FQ_RUN_ACCUMULATE as u32 } else { 0 };
        self.with_source(src, |s| check(self.ctx.raw, unsafe { sys::fq_pipe_launch_groupby(self.ctx.raw, self.raw, s, flags, stream) }))
    }

    /// Ok(Some(groups)), or Ok(None) when the table was too small (reserve more and relaunch)
    pub fn fetch_groupby(&self) -> FuseQueryResult<Option<u64>> {
        let mut n = 0u64;
        let st = unsafe { sys::fq_pipe_fetch_groupby(self.ctx.raw, self.raw, &mut n) };
        if st == sys::FQ_ERR_CAPACITY {
            return Ok(None);
        }
        check(self.ctx.raw, st)?;
        Ok(Some(n))
    }

    /// -> (key columns, key validity, leaf columns, leaf validity), `groups` rows each, table order
    #[allow(clippy::type_complexity)]
    pub fn export_groups(&self, groups: u64, stream: *mut c_void) -> FuseQueryResult<(Vec<Column>, Vec<Option<Column>>, Vec<Column>, Vec<Option<Column>>)> {
        let n_leaves = self.aggregator_nodes()?.len();
        let (mut keys, mut kval, mut leaves, mut lval) = (vec![], vec![], vec![], vec![]);
        for j in 0..self.n_keys {
            let (mut t, mut nullable) = (sys::FQ_NULL, 0);
            check(self.ctx.raw, unsafe { sys::fq_pipe_key_dtype(self.ctx.raw, self.raw, j as i32, &mut t, &mut nullable) })?;
            keys.push(Column::alloc(&self.ctx, t, groups.max(1))?);
            kval.push(if nullable != 0 { Some(Column::alloc(&self.ctx, sys::FQ_BOOL, groups.max(1))?) } else { None });
        }
        for k in 0..n_leaves {
            let (mut t, mut nullable) = (sys::FQ_NULL, 0);
            check(self.ctx.raw, unsafe { sys::fq_pipe_leaf_dtype(self.ctx.raw, self.raw, k as i32, &mut t, &mut nullable) })?;
            leaves.push(Column::alloc(&self.ctx, t, groups.max(1))?);
            lval.push(if nullable != 0 { Some(Column::alloc(&self.ctx, sys::FQ_BOOL, groups.max(1))?) } else { None });
        }
        let raw = |cols: &Vec<Column>| cols.iter().map(|c| c.raw).collect::<Vec<_>>();
        let raw_opt = |cols: &Vec<Option<Column>>| cols.iter().map(|c| c.as_ref().map_or(ptr::null_mut(), |c| c.raw)).collect::<Vec<_>>();
        let (kc, kv, lc, lv) = (raw(&keys), raw_opt(&kval), raw(&leaves), raw_opt(&lval));
        check(self.ctx.raw, unsafe {
            sys::fq_pipe_export_groups(self.ctx.raw, self.raw, kc.as_ptr(), kv.as_ptr(), lc.as_ptr(), lv.as_ptr(), groups, stream)
        })?;
        self.ctx.synchronize(stream)?;
        Ok((keys, kval, leaves, lval))
    }
}

impl Drop for Pipe {
    fn drop(&mut self) {
        unsafe { sys::fq_pipe_destroy(self.ctx.raw, self.raw) }
    }
}

/// A prepared statement: the launches issued between `Graph::record`'s begin and end, replayed with one graph launch
/// (fq_graph_*).  The pipes launched inside must outlive it; their fetch_* calls work after every replay as after a direct launch.
pub struct Graph {
    ctx: Arc<GpuContext>,
    raw: *mut sys::fq_graph,
}

unsafe impl Send for Graph {}

impl Graph {
    /// Records what `launches` issues on `stream` (an explicit stream; nothing runs) and instantiates it.
    pub fn record(ctx: &Arc<GpuContext>, stream: *mut c_void, launches: impl FnOnce() -> FuseQueryResult<()>) -> FuseQueryResult<Self> {
        check(ctx.raw, unsafe { sys::fq_graph_begin(ctx.raw, stream) })?;
        let issued = launches();
        let mut raw = ptr::null_mut();
        let ended = check(ctx.raw, unsafe { sys::fq_graph_end(ctx.raw, stream, &mut raw) });
        if let Err(e) = issued {
            if !raw.is_null() {
                unsafe { sys::fq_graph_destroy(ctx.raw, raw) }
            }
            return Err(e);
        }
        ended?;
        Ok(Graph { ctx: ctx.clone(), raw })
    }

    pub fn launch(&self, stream: *mut c_void) -> FuseQueryResult<()> {
        check(self.ctx.raw, unsafe { sys::fq_graph_launch(self.ctx.raw, self.raw, stream) })
    }
}

impl Drop for Graph {
    fn drop(&mut self) {
        unsafe { sys::fq_graph_destroy(self.ctx.raw, self.raw) }
    }
}
