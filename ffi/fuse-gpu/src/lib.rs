//! Safe wrappers over `fuse-gpu-sys` (the C ABI of libfuse_gpu.so, include/fuse_gpu.h) on the reference's own surface.
//!
//! UNCOMPILED: no Rust toolchain exists in the build image.  Written against fuse-query at the surveyed commit:
//! `FuseQueryError` (src/error.rs:10-28), `DataValue` (src/datavalues/data_value.rs:19-35), `ExpressionPlan`
//! (src/planners/plan_expression.rs:13-27), `IProcessor` (src/processors/processor.rs:22-58), `ITable`
//! (src/datasources/table.rs:13-22).  The C++ host mirror (fuse_query_b200/csrc/host) has the same structure and IS
//! compiled and tested; tests/abi_sequence.c replays this crate's call sequence in C.
//!
//! Layout
//!   lib.rs        GpuContext, Column, error mapping (`check`)
//!   pipe.rs       ExpressionPlan -> fq_expr_node[] lowering, Pipe (aggregate / projection / group-by launches)
//!   group.rs      Group: the cross-GPU merge point over peer memory
//!   transform.rs  GpuPipeTransform / GpuGroupByTransform : IProcessor
//!   table.rs      GpuNumbersTable : ITable (device-resident shards of system.numbers_mt)
#![allow(clippy::missing_safety_doc)]

pub mod group;
pub mod pipe;
pub mod table;
pub mod transform;

use std::ffi::CStr;
use std::os::raw::c_void;
use std::ptr;
use std::sync::Arc;

use fuse_gpu_sys as sys;

use crate::datavalues::{DataType, DataValue};
use crate::error::{FuseQueryError, FuseQueryResult};

/// Status + `fq_last_error` -> `FuseQueryError`.  The library's messages already carry the Display text of the
/// reference's errors ("Internal Error: ...", "Error during plan: ..."), so the prefix is stripped and the variant kept.
pub(crate) fn check(ctx: *const sys::fq_ctx, st: sys::fq_status) -> FuseQueryResult<()> {
    if st == sys::FQ_OK {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(sys::fq_last_error(ctx)) }.to_string_lossy().into_owned();
    let strip = |prefix: &str| msg.strip_prefix(prefix).unwrap_or(&msg).to_string();
    Err(if st == sys::FQ_ERR_PLAN {
        FuseQueryError::Plan(strip("Error during plan: "))
    } else {
        FuseQueryError::Internal(strip("Internal Error: "))
    })
}

/// `DataType` <-> `fq_dtype` (tags in the declaration order of DataValue, data_value.rs:19-35).
pub(crate) fn dtype_tag(t: &DataType) -> FuseQueryResult<sys::fq_dtype> {
    Ok(match t {
        DataType::Boolean => sys::FQ_BOOL,
        DataType::Int8 => sys::FQ_I8,
        DataType::Int16 => sys::FQ_I16,
        DataType::Int32 => sys::FQ_I32,
        DataType::Int64 => sys::FQ_I64,
        DataType::UInt8 => sys::FQ_U8,
        DataType::UInt16 => sys::FQ_U16,
        DataType::UInt32 => sys::FQ_U32,
        DataType::UInt64 => sys::FQ_U64,
        DataType::Float32 => sys::FQ_F32,
        DataType::Float64 => sys::FQ_F64,
        other => {
            return Err(FuseQueryError::Internal(format!(
                "Unsupported on the device path: column of type {:?}",
                other
            )))
        }
    })
}

pub(crate) fn dtype_of_tag(t: sys::fq_dtype) -> DataType {
    match t {
        sys::FQ_BOOL => DataType::Boolean,
        sys::FQ_I8 => DataType::Int8,
        sys::FQ_I16 => DataType::Int16,
        sys::FQ_I32 => DataType::Int32,
        sys::FQ_I64 => DataType::Int64,
        sys::FQ_U8 => DataType::UInt8,
        sys::FQ_U16 => DataType::UInt16,
        sys::FQ_U32 => DataType::UInt32,
        sys::FQ_U64 => DataType::UInt64,
        sys::FQ_F32 => DataType::Float32,
        sys::FQ_F64 => DataType::Float64,
        _ => DataType::Null,
    }
}

pub(crate) fn dtype_size(t: sys::fq_dtype) -> usize {
    match t {
        sys::FQ_BOOL | sys::FQ_I8 | sys::FQ_U8 => 1,
        sys::FQ_I16 | sys::FQ_U16 => 2,
        sys::FQ_I32 | sys::FQ_U32 | sys::FQ_F32 => 4,
        _ => 8,
    }
}

/// `fq_value` -> `DataValue` (what `Function::accumulate_result` holds per Aggregator leaf).
pub(crate) fn value_of(v: &sys::fq_value) -> DataValue {
    let some = v.some != 0;
    unsafe {
        match v.dtype {
            sys::FQ_NULL => DataValue::Null,
            sys::FQ_BOOL => DataValue::Boolean(if some { Some(v.v.i != 0) } else { None }),
            sys::FQ_I8 => DataValue::Int8(if some { Some(v.v.i as i8) } else { None }),
            sys::FQ_I16 => DataValue::Int16(if some { Some(v.v.i as i16) } else { None }),
            sys::FQ_I32 => DataValue::Int32(if some { Some(v.v.i as i32) } else { None }),
            sys::FQ_I64 => DataValue::Int64(if some { Some(v.v.i) } else { None }),
            sys::FQ_U8 => DataValue::UInt8(if some { Some(v.v.u as u8) } else { None }),
            sys::FQ_U16 => DataValue::UInt16(if some { Some(v.v.u as u16) } else { None }),
            sys::FQ_U32 => DataValue::UInt32(if some { Some(v.v.u as u32) } else { None }),
            sys::FQ_U64 => DataValue::UInt64(if some { Some(v.v.u) } else { None }),
            sys::FQ_F32 => DataValue::Float32(if some { Some(v.v.f as f32) } else { None }),
            sys::FQ_F64 => DataValue::Float64(if some { Some(v.v.f) } else { None }),
            _ => DataValue::Null,
        }
    }
}

/// One `fq_ctx` = one CUDA device.  Thread-safe on the library side; cloned freely behind an `Arc`.
pub struct GpuContext {
    pub(crate) raw: *mut sys::fq_ctx,
    pub device: i32,
}

unsafe impl Send for GpuContext {}
unsafe impl Sync for GpuContext {}

impl GpuContext {
    pub fn try_create(device: i32) -> FuseQueryResult<Arc<Self>> {
        if unsafe { sys::fq_abi_version() } != sys::FQ_ABI_VERSION as u32 {
            return Err(FuseQueryError::Internal("libfuse_gpu.so has another ABI version than these bindings".to_string()));
        }
        let mut raw = ptr::null_mut();
        // a failed create reports through fq_last_error(NULL); there is no CPU fallback behind it
        check(ptr::null(), unsafe { sys::fq_ctx_create(device, &mut raw) })?;
        Ok(Arc::new(GpuContext { raw, device }))
    }

    pub fn sm_count(&self) -> i32 {
        unsafe { sys::fq_ctx_sm_count(self.raw) }
    }

    pub fn launch_count(&self) -> u64 {
        unsafe { sys::fq_ctx_launch_count(self.raw) }
    }

    pub fn synchronize(&self, stream: *mut c_void) -> FuseQueryResult<()> {
        check(self.raw, unsafe { sys::fq_stream_synchronize(self.raw, stream) })
    }

    /// A non-blocking stream on this context's device (the crate has no CUDA bindings of its own); pair with `stream_destroy`.
    pub fn stream_create(&self) -> FuseQueryResult<*mut c_void> {
        let mut s = ptr::null_mut();
        check(self.raw, unsafe { sys::fq_stream_create(self.raw, &mut s) })?;
        Ok(s)
    }

    pub fn stream_destroy(&self, stream: *mut c_void) {
        unsafe { sys::fq_stream_destroy(self.raw, stream) }
    }
}

impl Drop for GpuContext {
    fn drop(&mut self) {
        unsafe { sys::fq_ctx_destroy(self.raw) }
    }
}

/// A device-resident Arrow-layout values buffer (`fq_column`), optionally with a validity column.
pub struct Column {
    pub(crate) ctx: Arc<GpuContext>,
    pub(crate) raw: *mut sys::fq_column,
    validity: Option<Box<Column>>,
    _parent: Option<Arc<Column>>,
}

unsafe impl Send for Column {}
unsafe impl Sync for Column {}

impl Column {
    pub fn alloc(ctx: &Arc<GpuContext>, dtype: sys::fq_dtype, len: u64) -> FuseQueryResult<Self> {
        let mut raw = ptr::null_mut();
        check(ctx.raw, unsafe { sys::fq_column_alloc(ctx.raw, dtype, len, &mut raw) })?;
        Ok(Column { ctx: ctx.clone(), raw, validity: None, _parent: None })
    }

    /// `system.numbers_mt` shard [begin, begin + n): one fill kernel (NumbersStream::poll_next, numbers_stream.rs:68-83).
    pub fn numbers(ctx: &Arc<GpuContext>, begin: u64, n: u64, stream: *mut c_void) -> FuseQueryResult<Self> {
        let col = Column::alloc(ctx, sys::FQ_U64, n)?;
        check(ctx.raw, unsafe { sys::fq_numbers_fill(ctx.raw, col.raw, 0, begin, n, stream) })?;
        Ok(col)
    }

    /// Upload an Arrow primitive array's values buffer (and its null bitmap, expanded on the device).
    pub fn from_arrow(ctx: &Arc<GpuContext>, array: &arrow::array::ArrayRef, stream: *mut c_void) -> FuseQueryResult<Self> {
        let dtype = dtype_tag(array.data_type())?;
        let data = array.data();
        let len = array.len() as u64;
        let mut col = Column::alloc(ctx, dtype, len)?;
        if dtype == sys::FQ_BOOL {
            // BooleanArray values are an LSB-first bitmap
            let bits = data.buffers()[0].raw_data();
            check(ctx.raw, unsafe {
                sys::fq_column_upload_bits(ctx.raw, col.raw, 0, bits as *const c_void, data.offset() as u64, len, stream)
            })?;
        } else {
            let bytes = unsafe { data.buffers()[0].raw_data().add(data.offset() * dtype_size(dtype)) };
            check(ctx.raw, unsafe { sys::fq_column_upload(ctx.raw, col.raw, 0, bytes as *const c_void, len, stream) })?;
        }
        if let Some(bitmap) = data.null_bitmap() {
            // arrow's validity buffer goes over as it is and STAYS bit-packed: pipes compiled with col_nullable = 2 read the
            // bits in place (1 bit of validity traffic per row); the array offset becomes the bit offset
            let bytes = bitmap.buffer_ref().len() as u64;
            let v = Column::alloc(ctx, sys::FQ_U8, bytes)?;
            check(ctx.raw, unsafe { sys::fq_column_upload(ctx.raw, v.raw, 0, bitmap.buffer_ref().raw_data() as *const c_void, bytes, stream) })?;
            check(ctx.raw, unsafe { sys::fq_column_set_validity_bitmap(ctx.raw, col.raw, v.raw, data.offset() as u64) })?;
            col.validity = Some(Box::new(v));
        }
        ctx.synchronize(stream)?;
        Ok(col)
    }

    /// attach a byte-per-row validity column this column does not own (the caller keeps it alive)
    pub fn set_validity(&self, validity: &Column) -> FuseQueryResult<()> {
        check(self.ctx.raw, unsafe { sys::fq_column_set_validity(self.ctx.raw, self.raw, validity.raw) })
    }

    /// the validity column uploaded or gathered with this column, if any
    pub fn validity(&self) -> Option<&Column> {
        self.validity.as_deref()
    }

    pub fn is_nullable(&self) -> bool {
        self.validity.is_some()
    }

    pub fn len(&self) -> u64 {
        unsafe { sys::fq_column_len(self.raw) }
    }

    pub fn is_empty(&self) -> bool {
        self.len() == 0
    }

    pub fn dtype(&self) -> sys::fq_dtype {
        unsafe { sys::fq_column_dtype(self.raw) }
    }

    /// ORDER BY: the row indexes (UInt32) that put `keys` in order — lexicographic, keys[0] most significant, ascending unless
    /// `descending[j]`, NULLs first, ties in input order (fq_sort_indices; the reference has no sort, README.md:28).
    pub fn sort_indices(ctx: &Arc<GpuContext>, keys: &[&Column], descending: &[bool], n_rows: u64, stream: *mut c_void) -> FuseQueryResult<Column> {
        let out = Column::alloc(ctx, sys::FQ_U32, n_rows.max(1))?;
        let raw: Vec<*const sys::fq_column> = keys.iter().map(|k| k.raw as *const sys::fq_column).collect();
        let desc: Vec<u8> = (0..keys.len()).map(|j| descending.get(j).copied().unwrap_or(false) as u8).collect();
        check(ctx.raw, unsafe { sys::fq_sort_indices(ctx.raw, raw.as_ptr(), desc.as_ptr(), raw.len() as i32, n_rows, out.raw, stream) })?;
        Ok(out)
    }

    /// ORDER BY ... LIMIT: the first min(limit, n_rows) indexes of the same order (radix select when one NOT NULL key decides).
    pub fn sort_indices_limit(ctx: &Arc<GpuContext>, keys: &[&Column], descending: &[bool], n_rows: u64, limit: u64, stream: *mut c_void)
                              -> FuseQueryResult<(Column, u64)> {
        let out = Column::alloc(ctx, sys::FQ_U32, limit.min(n_rows).max(1))?;
        let raw: Vec<*const sys::fq_column> = keys.iter().map(|k| k.raw as *const sys::fq_column).collect();
        let desc: Vec<u8> = (0..keys.len()).map(|j| descending.get(j).copied().unwrap_or(false) as u8).collect();
        let mut count = 0u64;
        check(ctx.raw, unsafe { sys::fq_sort_indices_limit(ctx.raw, raw.as_ptr(), desc.as_ptr(), raw.len() as i32, n_rows, limit, out.raw, &mut count, stream) })?;
        Ok((out, count))
    }

    /// out[i] = self[rows[i]] for the first `n` row indexes; the validity (bytes or bitmap) travels into a byte validity column.
    pub fn take(&self, rows: &Column, n: u64, nullable: bool, stream: *mut c_void) -> FuseQueryResult<Column> {
        let mut out = Column::alloc(&self.ctx, self.dtype(), n.max(1))?;
        let valid = if nullable { Some(Box::new(Column::alloc(&self.ctx, sys::FQ_BOOL, n.max(1))?)) } else { None };
        let vraw = valid.as_ref().map(|v| v.raw).unwrap_or(ptr::null_mut());
        check(self.ctx.raw, unsafe { sys::fq_column_take(self.ctx.raw, self.raw, rows.raw, n, out.raw, vraw, stream) })?;
        if let Some(v) = valid {
            check(self.ctx.raw, unsafe { sys::fq_column_set_validity(self.ctx.raw, out.raw, v.raw) })?;
            out.validity = Some(v);
        }
        Ok(out)
    }

    pub fn slice(self: &Arc<Self>, offset: u64, len: u64) -> FuseQueryResult<Column> {
        let mut raw = ptr::null_mut();
        check(self.ctx.raw, unsafe { sys::fq_column_slice(self.ctx.raw, self.raw, offset, len, &mut raw) })?;
        Ok(Column { ctx: self.ctx.clone(), raw, validity: None, _parent: Some(self.clone()) })
    }

    /// The first `n` rows as an Arrow array (values downloaded into a buffer the array then owns; validity packed into
    /// an Arrow bitmap by the device).
    pub fn to_arrow(&self, n: u64, validity: Option<&Column>, stream: *mut c_void) -> FuseQueryResult<arrow::array::ArrayRef> {
        use arrow::array::{make_array, ArrayData};
        use arrow::buffer::MutableBuffer;
        let dtype = self.dtype();
        let mut builder = ArrayData::builder(dtype_of_tag(dtype)).len(n as usize);
        if dtype == sys::FQ_BOOL {
            let mut bits = MutableBuffer::new(((n + 7) / 8) as usize).with_bitset(((n + 7) / 8) as usize, false);
            check(self.ctx.raw, unsafe {
                sys::fq_column_download_bits(self.ctx.raw, self.raw, 0, bits.raw_data_mut() as *mut c_void, n, stream)
            })?;
            self.ctx.synchronize(stream)?;
            builder = builder.add_buffer(bits.freeze());
        } else {
            let bytes = n as usize * dtype_size(dtype);
            let mut values = MutableBuffer::new(bytes);
            values.resize(bytes)?;
            check(self.ctx.raw, unsafe { sys::fq_column_download(self.ctx.raw, self.raw, 0, values.raw_data_mut() as *mut c_void, n, stream) })?;
            self.ctx.synchronize(stream)?;
            builder = builder.add_buffer(values.freeze());
        }
        if let Some(v) = validity {
            let mut bits = MutableBuffer::new(((n + 7) / 8) as usize).with_bitset(((n + 7) / 8) as usize, false);
            check(self.ctx.raw, unsafe {
                sys::fq_column_download_bits(self.ctx.raw, v.raw, 0, bits.raw_data_mut() as *mut c_void, n, stream)
            })?;
            self.ctx.synchronize(stream)?;
            builder = builder.null_bit_buffer(bits.freeze());
        }
        Ok(make_array(builder.build()))
    }
}

impl Drop for Column {
    fn drop(&mut self) {
        unsafe { sys::fq_column_free(self.ctx.raw, self.raw) }
    }
}
