//! Page-locked `Vec`s: `Vec<u64, PinnedAlloc>` backed by pmt_host_alloc (cudaHostAlloc, portable).
//!
//! pmt_merkle_tree_build / pmt_mmr_extend pipeline H2D copies, hashing and D2H copies over chunks; from PAGEABLE memory
//! the CUDA driver stages every copy through its own bounce buffer and the pipeline serialises (measured on B200 for
//! 2^24 x 4 leaves: 21.6 ms from pinned buffers, 114.6 ms from a plain Vec).  Allocate the big flat buffers
//! (`flat`, `dig` of wrappers.rs) with this allocator, or register existing ones once with pmt_host_register (23.1 ms per
//! build after a one-off ~180 ms for 1.5 GiB).
use std::alloc::{AllocError, Allocator, Layout};
use std::os::raw::c_void;
use std::ptr::NonNull;

use crate::pmt_ffi::*;

#[derive(Clone, Copy)]
pub struct PinnedAlloc {
    pub ctx: *mut pmt_ctx,
}

unsafe impl Allocator for PinnedAlloc {
    fn allocate(&self, layout: Layout) -> Result<NonNull<[u8]>, AllocError> {
        if layout.align() > 256 {
            return Err(AllocError);
        }
        let mut p: *mut c_void = std::ptr::null_mut();
        let rc = unsafe { pmt_host_alloc(self.ctx, layout.size().max(1), &mut p) };
        if rc != 0 || p.is_null() {
            return Err(AllocError);
        }
        Ok(NonNull::slice_from_raw_parts(unsafe { NonNull::new_unchecked(p as *mut u8) }, layout.size()))
    }
    unsafe fn deallocate(&self, ptr: NonNull<u8>, _layout: Layout) {
        pmt_host_free(self.ctx, ptr.as_ptr() as *mut c_void);
    }
}

/// a zeroed pinned Vec<u64> of `len` words
pub fn pinned_u64(ctx: &Ctx, len: usize) -> Vec<u64, PinnedAlloc> {
    let mut v = Vec::with_capacity_in(len, PinnedAlloc { ctx: ctx.0 });
    v.resize(len, 0u64);
    v
}
