//! `GpuNumbersTable: ITable` — system.numbers_mt whose partitions live in HBM (datasources/system/numbers_table.rs,
//! numbers_stream.rs).  read_plan is the reference's; what changes is what a partition *is*: a contiguous row range of a
//! device-resident UInt64 shard (materialised once by one fill kernel and kept across queries — they ARE the table) or,
//! in generated mode, nothing at all (the kernels compute number = begin + row).  UNCOMPILED (see lib.rs).

use std::collections::VecDeque;
use std::os::raw::c_void;
use std::sync::{Arc, Mutex};

use async_trait::async_trait;

use crate::datablocks::DataBlock;
use crate::datasources::system::NumbersTable;
use crate::datasources::{ITable, Partition};
use crate::datastreams::{DataBlockStream, SendableDataBlockStream};
use crate::datavalues::DataSchemaRef;
use crate::error::{FuseQueryError, FuseQueryResult};
use crate::planners::{PlanNode, ReadDataSourcePlan};

use super::{Column, GpuContext};

/// Rows a list of partitions emits, in NumbersStream::create order (numbers_stream.rs:27-62).  `tail_quirk` reproduces
/// :44-46: a partition of >= 10 000 rows that is not a multiple of 10 000 emits only 10 000 * (blocks - 1) + remain + 1 rows.
#[derive(Clone, Copy, Debug, PartialEq)]
pub struct RowRange {
    pub begin: u64,
    pub rows: u64,
}

pub fn emitted_ranges(parts: &[Partition], tail_quirk: bool, align_runs: bool) -> FuseQueryResult<Vec<RowRange>> {
    const BLOCK: u64 = 10_000;
    let mut out: Vec<RowRange> = vec![];
    for part in parts {
        let names: Vec<&str> = part.name.split('-').collect();
        if names.len() != 3 {
            return Err(FuseQueryError::Internal(format!("bad partition name {}", part.name)));
        }
        let begin: u64 = names[1].parse()?;
        let end: u64 = names[2].parse()?;
        let count = end - begin + 1;
        let (blocks, remain) = (count / BLOCK, count % BLOCK);
        let rows = if tail_quirk && blocks > 0 && remain > 0 { BLOCK * (blocks - 1) + remain + 1 } else { count };
        match out.last_mut() {
            // merging two partitions into one run keeps the reference's block boundaries only if the run so far is whole blocks
            Some(last) if last.begin + last.rows == begin && (!align_runs || last.rows % BLOCK == 0) => last.rows += rows,
            _ => out.push(RowRange { begin, rows }),
        }
    }
    Ok(out)
}

/// Shards stay resident across queries, least recently used first out (cap: 140 GB of the 180 GB of HBM).
pub struct ShardCache {
    items: Mutex<VecDeque<(RowRange, Arc<Column>)>>,
}

impl ShardCache {
    pub fn new() -> Self {
        ShardCache { items: Mutex::new(VecDeque::new()) }
    }

    pub fn get(&self, gpu: &Arc<GpuContext>, r: RowRange, stream: *mut c_void) -> FuseQueryResult<Arc<Column>> {
        let mut items = self.items.lock().map_err(|e| FuseQueryError::Internal(e.to_string()))?;
        if let Some(pos) = items.iter().position(|(c, _)| c.begin <= r.begin && r.begin + r.rows <= c.begin + c.rows && (r.begin - c.begin) % 2 == 0) {
            let (range, col) = items.remove(pos).unwrap();
            items.push_front((range, col.clone()));
            return Ok(if range == r { col } else { Arc::new(col.slice(r.begin - range.begin, r.rows)?) });
        }
        const CAP_BYTES: u64 = 140 << 30;
        while items.iter().map(|(c, _)| c.rows * 8).sum::<u64>() + r.rows * 8 > CAP_BYTES && !items.is_empty() {
            items.pop_back();
        }
        let col = Arc::new(Column::numbers(gpu, r.begin, r.rows, stream)?);
        items.push_front((r, col.clone()));
        Ok(col)
    }
}

impl Default for ShardCache {
    fn default() -> Self {
        Self::new()
    }
}

pub struct GpuNumbersTable {
    inner: NumbersTable,
    gpu: Arc<GpuContext>,
    pub shards: Arc<ShardCache>,
    pub tail_quirk: bool,
}

impl GpuNumbersTable {
    pub fn create(gpu: Arc<GpuContext>) -> Self {
        GpuNumbersTable { inner: NumbersTable::create(), gpu, shards: Arc::new(ShardCache::new()), tail_quirk: true }
    }

    /// The device-resident runs of `parts`: what a fused pipe launches over (one launch per run).
    pub fn device_runs(&self, parts: &[Partition], align_runs: bool, stream: *mut c_void) -> FuseQueryResult<Vec<(RowRange, Arc<Column>)>> {
        emitted_ranges(parts, self.tail_quirk, align_runs)?
            .into_iter()
            .map(|r| Ok((r, self.shards.get(&self.gpu, r, stream)?)))
            .collect()
    }
}

#[async_trait]
impl ITable for GpuNumbersTable {
    fn name(&self) -> &str {
        self.inner.name()
    }

    fn schema(&self) -> FuseQueryResult<DataSchemaRef> {
        self.inner.schema()
    }

    fn read_plan(&self, push_down_plan: PlanNode) -> FuseQueryResult<ReadDataSourcePlan> {
        self.inner.read_plan(push_down_plan) // partitioning is the reference's (numbers_table.rs:29-55)
    }

    /// Host-visible blocks for processors that are NOT fused (reference-shaped pipelines): each run comes back as Arrow
    /// arrays of 10 000 rows, downloaded from the resident shard.  Fused pipes never call this: they use `device_runs`.
    async fn read(&self, parts: Vec<Partition>) -> FuseQueryResult<SendableDataBlockStream> {
        let schema = self.inner.schema()?;
        let mut blocks = vec![];
        for (range, col) in self.device_runs(&parts, false, std::ptr::null_mut())? {
            let mut off = 0;
            while off < range.rows {
                let n = (range.rows - off).min(10_000);
                let piece = col.slice(off, n)?;
                blocks.push(DataBlock::create(schema.clone(), vec![piece.to_arrow(n, None, std::ptr::null_mut())?]));
                off += n;
            }
        }
        Ok(Box::pin(DataBlockStream::create(schema, None, blocks)))
    }
}
