//! The merge point across GPUs (processors/processor_merge.rs:37-66) over NVLink peer memory: one process per GPU, each
//! owning an exchange window.  UNCOMPILED (see lib.rs).

use std::os::raw::c_void;
use std::ptr;
use std::sync::Arc;

use fuse_gpu_sys as sys;

use crate::error::{FuseQueryError, FuseQueryResult};

use super::pipe::Pipe;
use super::{check, Column, GpuContext};

pub struct Group {
    ctx: Arc<GpuContext>,
    raw: *mut sys::fq_group,
    pub rank: i32,
    pub world: i32,
}

unsafe impl Send for Group {}

impl Group {
    pub fn try_create(ctx: &Arc<GpuContext>, rank: i32, world: i32, row_bytes: u64) -> FuseQueryResult<Self> {
        let mut raw = ptr::null_mut();
        check(ctx.raw, unsafe { sys::fq_group_create(ctx.raw, rank, world, row_bytes, &mut raw) })?;
        Ok(Group { ctx: ctx.clone(), raw, rank, world })
    }

    /// 64-byte CUDA IPC handle of this rank's window: send it to every peer (any transport).
    pub fn handle(&self) -> FuseQueryResult<[u8; 64]> {
        let mut h = [0u8; 64];
        check(self.ctx.raw, unsafe { sys::fq_group_handle(self.ctx.raw, self.raw, h.as_mut_ptr() as *mut c_void) })?;
        Ok(h)
    }

    /// `handles[r]` = rank r's handle (own entry ignored)
    pub fn connect(&mut self, handles: &[[u8; 64]]) -> FuseQueryResult<()> {
        if handles.len() != self.world as usize {
            return Err(FuseQueryError::Internal(format!("group of {} ranks got {} handles", self.world, handles.len())));
        }
        let blob: Vec<u8> = handles.iter().flat_map(|h| h.iter().copied()).collect();
        check(self.ctx.raw, unsafe { sys::fq_group_connect(self.ctx.raw, self.raw, blob.as_ptr() as *const c_void) })
    }

    /// Aggregate launches of `pipe` now end with the in-kernel exchange + fold; `Pipe::fetch_merged` reads the result.
    pub fn attach(&self, pipe: &Pipe) -> FuseQueryResult<()> {
        check(self.ctx.raw, unsafe { sys::fq_pipe_set_group(self.ctx.raw, pipe.raw(), self.raw) })
    }

    pub fn detach(&self, pipe: &Pipe) -> FuseQueryResult<()> {
        check(self.ctx.raw, unsafe { sys::fq_pipe_set_group(self.ctx.raw, pipe.raw(), ptr::null_mut()) })
    }

    /// MergeProcessor + the LimitTransform after it for projection pipes: every rank's kept rows in rank order, cut at `limit`.
    #[allow(clippy::too_many_arguments)]
    pub fn gather_project(&self, pipe: &Pipe, local: &[Column], local_valid: &[Option<Column>], finals: &[Column], final_valid: &[Option<Column>],
                          limit: Option<usize>, stream: *mut c_void) -> FuseQueryResult<(u64, u64)> {
        let raw = |cols: &[Column]| cols.iter().map(|c| c.raw).collect::<Vec<_>>();
        let raw_opt = |cols: &[Option<Column>]| cols.iter().map(|c| c.as_ref().map_or(ptr::null_mut(), |c| c.raw)).collect::<Vec<_>>();
        let (l, lv, f, fv) = (raw(local), raw_opt(local_valid), raw(finals), raw_opt(final_valid));
        check(self.ctx.raw, unsafe {
            sys::fq_group_gather_project(self.ctx.raw, self.raw, pipe.raw(), l.as_ptr(), lv.as_ptr(), f.as_ptr(), fv.as_ptr(),
                                         limit.map_or(-1i64, |n| n as i64), stream)
        })?;
        let (mut selected, mut rows) = (0u64, 0u64);
        check(self.ctx.raw, unsafe { sys::fq_group_fetch_gather(self.ctx.raw, self.raw, &mut selected, &mut rows) })?;
        Ok((selected, rows))
    }
}

impl Drop for Group {
    fn drop(&mut self) {
        unsafe { sys::fq_group_destroy(self.ctx.raw, self.raw) }
    }
}
