//! pmt-shim: Rust bindings of libpmt (include/pmt.h) and wrappers that return the reference's own types.
//!
//! NOT COMPILED in the image this repository is built in (no Rust toolchain there); kept in sync with include/pmt.h by
//! hand.  `pmt_ffi` = the `extern "C"` declarations, `wrappers` = drop-ins for `MerkleTree::build`
//! (simple_merkle_tree.rs:28-51), the batch form of `MMR::add_leaf` (merkle_mountain_ranges.rs:89-120) and upstream
//! plonky2's `MerkleTree::new`, `pinned` = page-locked `Vec`s (feature "pinned").
#![cfg_attr(feature = "pinned", feature(allocator_api))]
pub mod pmt_ffi;
pub mod wrappers;
#[cfg(feature = "pinned")]
pub mod pinned;
