//! The reference's MMR tests (/root/reference/src/mmr/merkle_mountain_ranges.rs:278-374) restated against the GPU wrapper
//! `mmr_extend_gpu`: the index tables are the reference's own (pure index math, unchanged); the append tests compare the
//! engine's `elements` with the reference's sequential `add_leaf` loop element by element, and the reference's own
//! `get_proof` / `MMR_proof::verify` / `bagging_the_peaks` run unchanged on the MMR the engine built.
//! UNCOMPILED in this repository's image; run with `cargo test -- --test-threads=1` on a B200 box.
use plonky2::field::goldilocks_field::GoldilocksField as F;
use plonky2::field::types::Field;
use plonky2_merkle_trees::mmr::common::GOLDILOCKS_FIELD_ORDER;
use plonky2_merkle_trees::mmr::merkle_mountain_ranges::{get_heights_bitmap_for_mmr_size, get_mmr_index, MMR};
use pmt_shim::pmt_ffi::{pmt_mmr_index, Ctx};
use pmt_shim::wrappers::{mmr_extend_gpu, mmr_leaf_count};
use rand::Rng;

fn random_leaves(n: usize) -> Vec<F> {
    let mut rng = rand::thread_rng();
    (0..n).map(|_| F::from_canonical_u64(rng.gen_range(0..GOLDILOCKS_FIELD_ORDER))).collect()
}

#[test]
fn test_heights_bitmap() {      // :280-297, the reference's table
    for (size, bitmap) in [(1, 1), (3, 2), (4, 3), (7, 4), (10, 6), (15, 8), (22, 12), (25, 14), (26, 15), (31, 16), (32, 17), (34, 18),
                           (35, 19), (38, 20), (41, 22), (42, 23)] {
        assert!(get_heights_bitmap_for_mmr_size(size) == (bitmap, 0));
    }
}

#[test]
fn test_get_mmr_index() {       // :307-324, and libpmt's closed form 2 i - popcount(i) agrees
    for (normal, idx) in [(0, 0), (1, 1), (2, 3), (3, 4), (4, 7), (5, 8), (6, 10), (7, 11), (8, 15), (9, 16), (10, 18), (11, 19), (12, 22),
                          (13, 23), (14, 25), (15, 26)] {
        assert!(get_mmr_index(normal) == idx);
        assert!(unsafe { pmt_mmr_index(normal as usize) } == idx as usize);
    }
}

#[test]
fn test_mmr_add_leaf() {        // :327-340: 100 leaves; here also against the sequential reference loop, in three batches
    let ctx = Ctx::new(0);
    let leaves = random_leaves(100);
    let mut want = MMR::new();
    for l in &leaves {
        want.add_leaf(*l);
    }
    let mut mmr = MMR::new();
    mmr_extend_gpu(&ctx, &mut mmr, &leaves[..37]);
    mmr_extend_gpu(&ctx, &mut mmr, &leaves[37..38]);
    mmr_extend_gpu(&ctx, &mut mmr, &leaves[38..]);
    assert!(mmr.elements == want.elements);
    assert_eq!(mmr_leaf_count(mmr.elements.len()), 100);
}

#[test]
fn test_get_proof() {           // :343-373: 16 leaves, proof for leaf 4 (mmr index 7), the reference's own prover and verifier
    let ctx = Ctx::new(0);
    let leaves = random_leaves(16);
    let mut mmr = MMR::new();
    mmr_extend_gpu(&ctx, &mut mmr, &leaves);
    let (standard_index, leaf_index) = (4, 7);
    let proof = mmr.clone().get_proof(leaf_index);
    let root = mmr.clone().bagging_the_peaks();
    assert!(proof.verify(leaves[standard_index], root));
    let mut seq = MMR::new();
    for l in &leaves {
        seq.add_leaf(*l);
    }
    assert!(seq.elements == mmr.elements && seq.bagging_the_peaks() == root);
}
