"""plonky2_merkle_trees_b200 -- B200-native Poseidon-Goldilocks Merkle / MMR commitment engine.

Drop-in for the one data-parallel hot path of hashcloak/plonky2-merkle-trees (leaf hashing + level-by-level
two_to_one): the CUDA kernels live in csrc/, the C ABI in include/pmt.h (libpmt.so), and the modules here mirror the
reference's Rust interface for that path on top of the C ABI:

    simple_merkle_tree.MerkleTree / verify_merkle_proof      <- src/simple_merkle_tree/simple_merkle_tree.rs
    mmr.MMR / MMR_proof / get_mmr_index / ...                 <- src/mmr/merkle_mountain_ranges.rs
    merkle_tree.MerkleTree / verify_merkle_proof_to_cap      <- plonky2 hash/merkle_tree.rs (MerkleTree::new)
    hasher.two_to_one / hash_or_noop / hash_no_pad           <- PoseidonHash: Hasher
    sharded.build_sharded_tree                               <- subtree-sharded multi-GPU build (one process per GPU)

There is no CPU fallback: without libpmt.so and a CUDA device every compute call raises.
"""
from ._lib import Context, PmtError, default_context, exported_symbols, load  # noqa: F401

__all__ = ["Context", "PmtError", "default_context", "exported_symbols", "load"]
