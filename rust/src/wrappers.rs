//! Wrappers that return the reference's own types (all their fields are `pub`, so they can be built from FFI output).
//!
//! `HashOut<GoldilocksField>` is a struct over `[GoldilocksField; 4]` and `GoldilocksField` is `#[repr(transparent)]` over
//! `u64`, but `HashOut` is not `repr(C)`: digests are converted element-wise.  Inputs may be non-canonical (`x.0` is
//! passed as it is, like the reference's `from_noncanonical_u128` leaves); outputs of libpmt are canonical by contract.
use plonky2::field::goldilocks_field::GoldilocksField as F;
use plonky2::field::types::{Field, PrimeField64};
use plonky2::hash::hash_types::HashOut;
use plonky2::hash::merkle_tree::{MerkleCap, MerkleTree as Plonky2MerkleTree};
use plonky2::hash::poseidon::PoseidonHash;
use plonky2_merkle_trees::mmr::merkle_mountain_ranges::MMR;
use plonky2_merkle_trees::simple_merkle_tree::simple_merkle_tree::MerkleTree;

use crate::pmt_ffi::*;

#[inline]
fn digest(w: &[u64]) -> HashOut<F> {
    HashOut { elements: [F::from_canonical_u64(w[0]), F::from_canonical_u64(w[1]), F::from_canonical_u64(w[2]), F::from_canonical_u64(w[3])] }
}

/// drop-in for `MerkleTree::build` (simple_merkle_tree.rs:28-51)
pub fn build_gpu(ctx: &Ctx, leaves: Vec<F>) -> MerkleTree {
    let n = leaves.len();
    let raw: Vec<u64> = leaves.iter().map(|x| x.0).collect();
    let mut levels = vec![0u64; (2 * n).saturating_sub(2) * 4];
    let mut root = [0u64; 4];
    // n not a power of two / n < 2: PMT_E_NOT_POW2 / PMT_E_INVALID_ARG -> panic, like log2_strict (:30) and the underflow (:38)
    ctx.check(unsafe { pmt_simple_tree_build(ctx.0, raw.as_ptr(), n, levels.as_mut_ptr(), root.as_mut_ptr()) });
    let (mut tree, mut off, mut m) = (Vec::new(), 0usize, n);
    while m >= 2 {
        tree.push(levels[off * 4..(off + m) * 4].chunks_exact(4).map(digest).collect());   // level-major -> Vec<Vec<HashOut>>
        off += m;
        m /= 2;
    }
    MerkleTree { count_levels: n.trailing_zeros() as usize, tree, root: digest(&root) }
}

/// number of leaves of an MMR from its element count (size = 2n - popcount(n) is strictly increasing in n)
pub fn mmr_leaf_count(size: usize) -> usize {
    let (mut lo, mut hi) = (0usize, size + 1);
    while lo < hi {
        let mid = (lo + hi) / 2;
        if unsafe { pmt_mmr_size(mid) } < size { lo = mid + 1 } else { hi = mid }
    }
    assert!(unsafe { pmt_mmr_size(lo) } == size, "corrupt MMR length {size}");
    lo
}

/// batch form of `for l in leaves { mmr.add_leaf(l) }` (merkle_mountain_ranges.rs:89-120); resumable.
/// libpmt reads only the popcount(n_before) old PEAKS of `elements` and writes only the new elements, so only those are
/// converted: the cost is O(m + log n), like the add_leaf calls it replaces, whatever the size of the MMR.
pub fn mmr_extend_gpu(ctx: &Ctx, mmr: &mut MMR, new_leaves: &[F]) {
    if new_leaves.is_empty() {
        return;
    }
    let old = mmr.elements.len();
    let n_before = mmr_leaf_count(old);
    let total = unsafe { pmt_mmr_size(n_before + new_leaves.len()) };
    // a buffer positioned like `elements`; the prefix is left untouched except for the peaks
    let mut buf = vec![0u64; total * 4];
    let mut base = 0usize;
    for bit in (0..usize::BITS).rev() {
        if (n_before >> bit) & 1 == 1 {
            base += 1usize << bit;
            let pos = unsafe { pmt_mmr_size(base) } - 1;                      // get_peaks (:179-200)
            for j in 0..4 {
                buf[4 * pos + j] = mmr.elements[pos].elements[j].to_canonical_u64();
            }
        }
    }
    let raw: Vec<u64> = new_leaves.iter().map(|x| x.0).collect();
    ctx.check(unsafe { pmt_mmr_extend(ctx.0, buf.as_mut_ptr(), n_before, raw.as_ptr(), raw.len()) });
    mmr.elements.extend(buf[old * 4..].chunks_exact(4).map(digest));
}

/// the same over every GPU of the box from this one process (one Ctx per device)
pub fn mmr_extend_multi_gpu(ctxs: &[Ctx], mmr: &mut MMR, new_leaves: &[F]) {
    if new_leaves.is_empty() {
        return;
    }
    let old = mmr.elements.len();
    let n_before = mmr_leaf_count(old);
    let total = unsafe { pmt_mmr_size(n_before + new_leaves.len()) };
    let mut buf = vec![0u64; total * 4];
    for (i, h) in mmr.elements.iter().enumerate() {      // the multi form also reads the peaks at its block boundaries: keep it simple
        for j in 0..4 {
            buf[4 * i + j] = h.elements[j].to_canonical_u64();
        }
    }
    let raw: Vec<u64> = new_leaves.iter().map(|x| x.0).collect();
    let handles: Vec<*mut pmt_ctx> = ctxs.iter().map(|c| c.0).collect();
    ctxs[0].check(unsafe { pmt_mmr_extend_multi(handles.as_ptr(), handles.len(), buf.as_mut_ptr(), n_before, raw.as_ptr(), raw.len()) });
    mmr.elements.extend(buf[old * 4..].chunks_exact(4).map(digest));
}

/// drop-in for plonky2's `MerkleTree::new(leaves, cap_height)` (all three fields of upstream's struct are `pub`)
pub fn plonky2_tree_gpu(ctx: &Ctx, leaves: Vec<Vec<F>>, cap_height: usize) -> Plonky2MerkleTree<F, PoseidonHash> {
    let (n, w) = (leaves.len(), leaves[0].len());
    let flat: Vec<u64> = leaves.iter().flat_map(|r| r.iter().map(|x| x.0)).collect();
    let (ncap, ndig) = (1usize << cap_height, 2 * (n - (1usize << cap_height)));
    let (mut dig, mut cap) = (vec![0u64; ndig * 4], vec![0u64; ncap * 4]);
    ctx.check(unsafe { pmt_merkle_tree_build(ctx.0, flat.as_ptr(), n, w, cap_height as u32, dig.as_mut_ptr(), cap.as_mut_ptr()) });
    Plonky2MerkleTree {
        leaves,
        digests: dig.chunks_exact(4).map(digest).collect(),
        cap: MerkleCap(cap.chunks_exact(4).map(digest).collect()),
    }
}

/// `plonky2_tree_gpu` over every device: one Ctx per GPU, one host thread per Ctx inside the call
pub fn plonky2_tree_multi_gpu(ctxs: &[Ctx], leaves: Vec<Vec<F>>, cap_height: usize) -> Plonky2MerkleTree<F, PoseidonHash> {
    let (n, w) = (leaves.len(), leaves[0].len());
    let flat: Vec<u64> = leaves.iter().flat_map(|r| r.iter().map(|x| x.0)).collect();
    let (ncap, ndig) = (1usize << cap_height, 2 * (n - (1usize << cap_height)));
    let (mut dig, mut cap) = (vec![0u64; ndig * 4], vec![0u64; ncap * 4]);
    let handles: Vec<*mut pmt_ctx> = ctxs.iter().map(|c| c.0).collect();
    ctxs[0].check(unsafe {
        pmt_merkle_tree_build_multi(handles.as_ptr(), handles.len(), flat.as_ptr(), n, w, cap_height as u32, dig.as_mut_ptr(), cap.as_mut_ptr())
    });
    Plonky2MerkleTree {
        leaves,
        digests: dig.chunks_exact(4).map(digest).collect(),
        cap: MerkleCap(cap.chunks_exact(4).map(digest).collect()),
    }
}

// ---- one process per GPU (launched by the host's own launcher: MPI, a job scheduler ...) ------------------------------------

/// a device buffer of `words` u64 owned by a Ctx (pmt_malloc / pmt_free)
struct DevBuf<'a> {
    ctx: &'a Ctx,
    ptr: *mut u64,
    words: usize,
}
impl<'a> DevBuf<'a> {
    fn new(ctx: &'a Ctx, words: usize) -> Self {
        let mut p: *mut std::os::raw::c_void = std::ptr::null_mut();
        ctx.check(unsafe { pmt_malloc(ctx.0, words.max(1) * 8, &mut p) });
        DevBuf { ctx, ptr: p as *mut u64, words }
    }
    fn upload(&self, src: &[u64]) {
        assert!(src.len() <= self.words);
        self.ctx.check(unsafe { pmt_memcpy_h2d(self.ctx.0, self.ptr as *mut _, src.as_ptr() as *const _, src.len() * 8) });
    }
    fn download(&self, words: usize) -> Vec<u64> {
        let mut out = vec![0u64; words];
        self.ctx.check(unsafe { pmt_memcpy_d2h(self.ctx.0, out.as_mut_ptr() as *mut _, self.ptr as *const _, words * 8) });
        out
    }
}
impl<'a> Drop for DevBuf<'a> {
    fn drop(&mut self) {
        unsafe { pmt_free(self.ctx.0, self.ptr as *mut _) };
    }
}

/// rank 0 creates the id and hands the 128 bytes to the other ranks by the host's own means
pub fn nccl_unique_id(ctx: &Ctx) -> [u8; 128] {
    let mut id = [0u8; 128];
    ctx.check(unsafe { pmt_nccl_unique_id(ctx.0, id.as_mut_ptr() as *mut _) });
    id
}

/// collective over the `world` ranks (a power of two); afterwards `comm_uses_peer_memory` tells whether the subtree roots
/// travel through peer-memory mailboxes (k_exchange_top) or through ncclAllGather
pub fn comm_init(ctx: &Ctx, unique_id: &[u8; 128], rank: usize, world: usize) {
    ctx.check(unsafe { pmt_comm_init(ctx.0, unique_id.as_ptr() as *const _, rank as i32, world as i32) });
}
pub fn comm_uses_peer_memory(ctx: &Ctx) -> bool {
    unsafe { pmt_comm_uses_peer_memory(ctx.0) != 0 }
}

/// this rank's share of `MerkleTree::new(leaves, cap_height)` over `world` ranks: `local_digests` is the contiguous slice of
/// upstream's `digests` that belongs to the rank's leaves, `roots` / `top` the gathered subtree roots and the levels above
/// them (empty when cap_height >= log2 world), `cap` the whole cap -- the same on every rank
pub struct ShardedTree {
    pub local_digests: Vec<HashOut<F>>,
    pub roots: Vec<HashOut<F>>,
    pub top: Vec<HashOut<F>>,
    pub cap: MerkleCap<F, PoseidonHash>,
}

/// collective: every rank passes ITS n_total / world leaf rows (rank r owns rows [r n_total / world, (r + 1) n_total / world))
pub fn plonky2_tree_sharded(ctx: &Ctx, local_leaves: &[Vec<F>], n_total: usize, cap_height: usize, world: usize) -> ShardedTree {
    let (per, w) = (local_leaves.len(), local_leaves[0].len());
    assert_eq!(per * world, n_total);
    let g = world.trailing_zeros() as usize;
    let ncap = 1usize << cap_height;
    let local_cap = if cap_height >= g { ncap / world } else { 1 };
    let flat: Vec<u64> = local_leaves.iter().flat_map(|r| r.iter().map(|x| x.0)).collect();
    let d_leaves = DevBuf::new(ctx, per * w);
    d_leaves.upload(&flat);
    let ndig = 2 * (per - local_cap);
    let ntop = if cap_height >= g { 0 } else { world - ncap };
    let (d_dig, d_roots, d_top, d_cap) = (DevBuf::new(ctx, ndig * 4), DevBuf::new(ctx, world * 4), DevBuf::new(ctx, ntop * 4), DevBuf::new(ctx, ncap * 4));
    ctx.check(unsafe {
        pmt_merkle_tree_build_sharded_dev(ctx.0, d_leaves.ptr, n_total, w, cap_height as u32, d_dig.ptr, d_roots.ptr, d_top.ptr, d_cap.ptr)
    });
    ctx.check(unsafe { pmt_sync(ctx.0) });
    let conv = |v: Vec<u64>| -> Vec<HashOut<F>> { v.chunks_exact(4).map(digest).collect() };
    ShardedTree {
        local_digests: conv(d_dig.download(ndig * 4)),
        roots: if ntop > 0 { conv(d_roots.download(world * 4)) } else { Vec::new() },
        top: conv(d_top.download(ntop * 4)),
        cap: MerkleCap(conv(d_cap.download(ncap * 4))),
    }
}
