//! `extern "C"` declarations of include/pmt.h (the subset a host needs; the `*_dev` twins take device pointers).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct pmt_ctx {
    _private: [u8; 0],
}

pub const PMT_OK: c_int = 0;
pub const PMT_E_INVALID_ARG: c_int = -1;
pub const PMT_E_NOT_POW2: c_int = -2;
pub const PMT_E_OOM: c_int = -3;
pub const PMT_E_CUDA: c_int = -4;
pub const PMT_E_RANGE: c_int = -5;
pub const PMT_E_NCCL: c_int = -6;

extern "C" {
    // context
    pub fn pmt_init(out: *mut *mut pmt_ctx, device_id: c_int) -> c_int;
    pub fn pmt_destroy(ctx: *mut pmt_ctx);
    pub fn pmt_last_error(ctx: *const pmt_ctx) -> *const c_char;
    pub fn pmt_sync(ctx: *mut pmt_ctx) -> c_int;
    // page-locked host memory (the pipelined host-buffer builders overlap copies and hashing only from pinned buffers)
    pub fn pmt_host_register(ctx: *mut pmt_ctx, ptr: *mut c_void, bytes: usize) -> c_int;
    pub fn pmt_host_unregister(ctx: *mut pmt_ctx, ptr: *mut c_void) -> c_int;
    pub fn pmt_host_alloc(ctx: *mut pmt_ctx, bytes: usize, ptr_out: *mut *mut c_void) -> c_int;
    pub fn pmt_host_free(ctx: *mut pmt_ctx, ptr: *mut c_void) -> c_int;
    // device memory helpers
    pub fn pmt_malloc(ctx: *mut pmt_ctx, bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn pmt_free(ctx: *mut pmt_ctx, dptr: *mut c_void) -> c_int;
    pub fn pmt_memcpy_h2d(ctx: *mut pmt_ctx, dst_dev: *mut c_void, src_host: *const c_void, bytes: usize) -> c_int;
    pub fn pmt_memcpy_d2h(ctx: *mut pmt_ctx, dst_host: *mut c_void, src_dev: *const c_void, bytes: usize) -> c_int;
    // Hasher
    pub fn pmt_hash_two_to_one(ctx: *mut pmt_ctx, l: *const u64, r: *const u64, n: usize, out: *mut u64) -> c_int;
    pub fn pmt_hash_or_noop(ctx: *mut pmt_ctx, rows: *const u64, n_rows: usize, width: usize, out: *mut u64) -> c_int;
    pub fn pmt_hash_no_pad(ctx: *mut pmt_ctx, rows: *const u64, n_rows: usize, width: usize, out: *mut u64) -> c_int;
    // simple tree (simple_merkle_tree.rs)
    pub fn pmt_simple_tree_build(ctx: *mut pmt_ctx, leaves: *const u64, n: usize, levels_out: *mut u64, root_out: *mut u64) -> c_int;
    pub fn pmt_simple_tree_prove(ctx: *mut pmt_ctx, levels: *const u64, n: usize, idx: *const u64, n_idx: usize, siblings_out: *mut u64) -> c_int;
    pub fn pmt_simple_tree_verify(ctx: *mut pmt_ctx, leaves: *const u64, idx: *const u64, n_idx: usize, root: *const u64,
                                  proofs: *const u64, path_len: usize, ok_out: *mut u8) -> c_int;
    // upstream plonky2 MerkleTree::new
    pub fn pmt_merkle_tree_build(ctx: *mut pmt_ctx, leaves: *const u64, n: usize, width: usize, cap_height: u32,
                                 digests_out: *mut u64, cap_out: *mut u64) -> c_int;
    pub fn pmt_merkle_tree_build_multi(ctxs: *const *mut pmt_ctx, n_ctx: usize, leaves: *const u64, n: usize, width: usize,
                                       cap_height: u32, digests_out: *mut u64, cap_out: *mut u64) -> c_int;
    pub fn pmt_merkle_tree_build_multi_dev(ctxs: *const *mut pmt_ctx, n_ctx: usize, d_leaves: *const *const u64, n: usize, width: usize,
                                           cap_height: u32, d_digests: *const *mut u64, d_roots: *mut u64, d_top: *mut u64,
                                           d_cap: *mut u64) -> c_int;
    pub fn pmt_merkle_prove(ctx: *mut pmt_ctx, digests: *const u64, n: usize, cap_height: u32, idx: *const u64, n_idx: usize,
                            siblings_out: *mut u64) -> c_int;
    pub fn pmt_merkle_verify(ctx: *mut pmt_ctx, leaf_rows: *const u64, width: usize, idx: *const u64, n_idx: usize, cap: *const u64,
                             cap_height: u32, proofs: *const u64, path_len: usize, ok_out: *mut u8) -> c_int;
    // one process per GPU: NCCL inside the library
    pub fn pmt_nccl_unique_id(ctx: *mut pmt_ctx, id_out_128_bytes: *mut c_void) -> c_int;
    pub fn pmt_comm_init(ctx: *mut pmt_ctx, unique_id_128_bytes: *const c_void, rank: c_int, world: c_int) -> c_int;
    pub fn pmt_comm_destroy(ctx: *mut pmt_ctx) -> c_int;
    pub fn pmt_comm_uses_peer_memory(ctx: *const pmt_ctx) -> c_int;
    pub fn pmt_merkle_tree_build_sharded_dev(ctx: *mut pmt_ctx, d_local_leaves: *const u64, n_total: usize, width: usize, cap_height: u32,
                                             d_local_digests: *mut u64, d_roots: *mut u64, d_top: *mut u64, d_cap: *mut u64) -> c_int;
    pub fn pmt_mmr_shard_plan(n_total: usize, world: usize, n_rounds: *mut u32, m_out_64: *mut usize, tail: *mut usize) -> c_int;
    pub fn pmt_mmr_build_sharded_dev(ctx: *mut pmt_ctx, d_local_leaves: *const u64, n_total: usize, d_local_elements: *mut u64,
                                     d_tail_elements: *mut u64, d_gathered: *mut u64, d_tops: *mut u64, d_peaks: *mut u64) -> c_int;
    // MMR (merkle_mountain_ranges.rs)
    pub fn pmt_mmr_size(n_leaves: usize) -> usize;
    pub fn pmt_mmr_index(leaf_normal_index: usize) -> usize;
    pub fn pmt_mmr_extend(ctx: *mut pmt_ctx, elements: *mut u64, n_before: usize, new_leaves: *const u64, m: usize) -> c_int;
    pub fn pmt_mmr_extend_multi(ctxs: *const *mut pmt_ctx, n_ctx: usize, elements: *mut u64, n_before: usize, new_leaves: *const u64,
                                m: usize) -> c_int;
    pub fn pmt_mmr_bag(ctx: *mut pmt_ctx, elements: *const u64, n_leaves: usize, root_out: *mut u64) -> c_int;
    pub fn pmt_mmr_peaks(ctx: *mut pmt_ctx, elements: *const u64, n_leaves: usize, peaks_out: *mut u64, n_peaks_out: *mut u32) -> c_int;
    pub fn pmt_mmr_prove(ctx: *mut pmt_ctx, elements: *const u64, n_leaves: usize, leaf_idx: *const u64, n_idx: usize,
                         siblings_out: *mut u64, on_left_out: *mut u8, path_len_out: *mut u32) -> c_int;
    pub fn pmt_mmr_verify(ctx: *mut pmt_ctx, leaves: *const u64, n_idx: usize, siblings: *const u64, on_left: *const u8,
                          path_len: *const u32, peaks: *const u64, n_peaks: u32, root: *const u64, status_out: *mut i8) -> c_int;
    // the prover's feed (upstream PolynomialBatch): column-major LDE values on the device
    pub fn pmt_merkle_tree_build_from_columns_dev(ctx: *mut pmt_ctx, d_columns: *const u64, n: usize, width: usize, bit_reverse: c_int,
                                                  cap_height: u32, d_leaves_out: *mut u64, d_digests: *mut u64, d_cap: *mut u64) -> c_int;
}

/// One `pmt_ctx`: one CUDA device, one stream.  Not `Sync`: use one per host thread.
pub struct Ctx(pub *mut pmt_ctx);

impl Ctx {
    pub fn new(device: i32) -> Self {
        let mut p = std::ptr::null_mut();
        let rc = unsafe { pmt_init(&mut p, device) };
        assert!(rc == 0, "pmt_init failed ({rc}): no usable CUDA device -- libpmt has no CPU fallback");
        Ctx(p)
    }
    /// non-zero status -> panic, like the reference's own asserts / log2_strict (simple_merkle_tree.rs:30,56,77)
    pub fn check(&self, rc: c_int) {
        if rc != 0 {
            let msg = unsafe { std::ffi::CStr::from_ptr(pmt_last_error(self.0)) }.to_string_lossy().into_owned();
            panic!("libpmt error {rc}: {msg}");
        }
    }
}
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { pmt_destroy(self.0) }
    }
}
unsafe impl Send for Ctx {}
