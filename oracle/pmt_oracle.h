/*
 * pmt_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the reference's hot path: Poseidon-over-Goldilocks leaf hashing plus
 * level-by-level two_to_one compression, as used by
 *   - MerkleTree::build / get_merkle_proof / verify_merkle_proof
 *         /root/reference/src/simple_merkle_tree/simple_merkle_tree.rs:21-109
 *   - MMR::add_leaf / get_peaks / bagging_the_peaks / get_proof / MMR_proof::verify
 *         /root/reference/src/mmr/merkle_mountain_ranges.rs:39-270
 *   - plonky2's MerkleTree::new / prove / verify_merkle_proof_to_cap (un-vendored dependency
 *     plonky2 v0.1.3 @ 3b21b87d0ab3f8ef4b9ff0b9dd70f8e32f5573f4, Cargo.toml:7 / Cargo.lock:460-462:
 *     plonky2/src/hash/{poseidon.rs,poseidon_goldilocks.rs,hashing.rs,merkle_tree.rs,merkle_proofs.rs},
 *     plonky2/src/plonk/config.rs, field/src/goldilocks_field.rs) -- restated from the published algorithm.
 *
 * Parity pinning: checked against every known-answer vector the reference's own tests hold for this path
 * (simple_merkle_tree.rs:136-140, 181-190, 210-211; merkle_mountain_ranges.rs:280-297, 307-324) and the two
 * upstream permutation test vectors -- see tests/test_oracle_kat.py.  Multi-block sponge output (hash_no_pad with
 * more than 8 felts) and the upstream `digests` layout are NOT pinned by any reference test ("parity unpinned"
 * for those two items; they rest on the upstream specification).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this library.
 */
#ifndef PMT_ORACLE_H
#define PMT_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- field + permutation ------------------------------------------------------------------------ */
uint64_t pmt_oracle_canonical(uint64_t x);
uint64_t pmt_oracle_mul(uint64_t a, uint64_t b);
/* naive 30-round specification (add constants, x^7, circulant MDS) */
void pmt_oracle_permute(uint64_t state[12]);
/* the same permutation executed with upstream's "fast partial rounds" restructuring (CPU-baseline speed) */
void pmt_oracle_permute_fast(uint64_t state[12]);

/* ---- Hasher ------------------------------------------------------------------------------------- */
void pmt_oracle_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]);
void pmt_oracle_hash_no_pad(const uint64_t *in, size_t n, uint64_t out[4]);
void pmt_oracle_hash_or_noop(const uint64_t *in, size_t n, uint64_t out[4]);

/* ---- simple_merkle_tree.rs ---------------------------------------------------------------------- */
/* levels_out: (2n-2)*4 u64, level-major (level 0 = n leaf digests, ..., last level = 2 digests). 0 on success. */
int pmt_oracle_simple_tree_build(const uint64_t *leaves, size_t n, uint64_t *levels_out, uint64_t root_out[4]);
int pmt_oracle_simple_tree_proof(const uint64_t *levels, size_t n, size_t leaf_index, uint64_t *proof_out /* log2(n)*4 */);
int pmt_oracle_simple_tree_in_between(const uint64_t *levels, const uint64_t root[4], size_t n, size_t leaf_index,
                                      uint64_t *out /* log2(n)*4 */);
int pmt_oracle_simple_tree_verify(uint64_t leaf, size_t leaf_index, const uint64_t root[4], const uint64_t *hashes,
                                  size_t n_hashes);

/* ---- merkle_mountain_ranges.rs ------------------------------------------------------------------ */
void pmt_oracle_mmr_heights_bitmap(size_t mmr_size, uint64_t *peaks_out, size_t *rem_out);
size_t pmt_oracle_mmr_index(size_t leaf_normal_index);
/* elements: capacity for the grown MMR (4 u64 per element); *len = number of elements, updated. */
void pmt_oracle_mmr_add_leaf(uint64_t *elements, size_t *len, uint64_t leaf);
/* returns number of peaks written (4 u64 each) */
size_t pmt_oracle_mmr_peaks(const uint64_t *elements, size_t len, uint64_t *peaks_out);
void pmt_oracle_mmr_bag(const uint64_t *elements, size_t len, uint64_t root_out[4]);
/* returns the path length; siblings_out 4 u64 per entry, on_left_out 1 byte per entry */
size_t pmt_oracle_mmr_subtree_proof(const uint64_t *elements, size_t len, size_t mmr_index, uint64_t *siblings_out,
                                    uint8_t *on_left_out);
/* 1 = true, 0 = false, -1 = the reference would panic (assert at merkle_mountain_ranges.rs:245) */
int pmt_oracle_mmr_verify(uint64_t leaf, const uint64_t root[4], const uint64_t *siblings, const uint8_t *on_left,
                          size_t path_len, const uint64_t *peaks, size_t n_peaks);

/* ---- plonky2 MerkleTree::new / prove / verify_merkle_proof_to_cap -------------------------------- */
/* leaves row-major n x w; digests_out 2(n-2^h)*4 u64 in upstream's interleaved layout; cap_out 2^h*4 u64.
 * threads <= 1: sequential; otherwise an OpenMP restatement of rayon's par_chunks + join. use_fast selects
 * pmt_oracle_permute_fast (identical output). */
int pmt_oracle_merkle_tree_new(const uint64_t *leaves, size_t n, size_t w, unsigned cap_height, uint64_t *digests_out,
                               uint64_t *cap_out, int threads, int use_fast);
int pmt_oracle_merkle_prove(const uint64_t *digests, size_t n, unsigned cap_height, size_t leaf_index,
                            uint64_t *siblings_out /* (log2 n - h)*4 */);
int pmt_oracle_merkle_verify_to_cap(const uint64_t *leaf, size_t w, size_t leaf_index, const uint64_t *cap,
                                    unsigned cap_height, const uint64_t *siblings, size_t n_siblings);
int pmt_oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
