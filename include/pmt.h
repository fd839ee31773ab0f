/*
 * pmt.h -- C ABI of libpmt.so: B200-native Poseidon-Goldilocks Merkle commitment engine.
 *
 * The reference (hashcloak/plonky2-merkle-trees) is pure Rust with NO FFI of its own; this header is the boundary a
 * maintainer binds with `extern "C"` (see INTEGRATION.md for the Rust shim) so that the reference's public types keep
 * their shape.  Each entry point names the reference function it replaces (file:line under /root/reference, or
 * [UPSTREAM] = plonky2 v0.1.3 @ 3b21b87d, the un-vendored dependency of Cargo.toml:7).
 *
 * Conventions
 *   - felt   = one Goldilocks element as little-endian u64.  INPUTS may be any u64 (non-canonical allowed, like
 *              GoldilocksField.0); OUTPUTS are always canonical (< p = 2^64 - 2^32 + 1).
 *   - digest = HashOut = 4 consecutive felts (32 bytes).
 *   - plain pointers + sizes, caller owns every buffer; nothing returned by the library outlives the ctx.
 *   - return value: 0 = PMT_OK, negative = PMT_E_*; pmt_last_error(ctx) has the message.  Never unwinds.
 *   - a ctx is bound to one CUDA device and one stream and is NOT thread-safe; use one ctx per host thread / rank.
 *   - *_dev entry points take DEVICE pointers, only ENQUEUE work on the ctx stream (no host sync) and return; call
 *     pmt_sync() before reading results on the host.  The host-buffer entry points are synchronous.
 *   - there is no CPU fallback: every entry point that computes fails with PMT_E_CUDA when no device is usable.
 *   - tuning knobs read from the environment at pmt_init / call time (measurement aids; the defaults are the measured
 *     optima, results never depend on them): PMT_COOP_MAX_LOG2 (levels of at most 2^k nodes run four threads per node,
 *     default 13), PMT_FUSE_SUBTREES (0: one launch per small level instead of ONE launch for the whole tail of a tree),
 *     PMT_PIPELINE_LOG2_CHUNKS (chunks of the pipelined host-buffer tree build, default 4), PMT_WAVE (0: one launch per big
 *     level instead of ONE wavefront launch for all of them), PMT_WAVE_MIN_LOG2 (levels of fewer than 2^k nodes stay out of
 *     the wavefront, default 14), PMT_COPY_THREADS (host threads that stage pageable caller buffers), PMT_EXCHANGE (nccl:
 *     sharded builds exchange their roots with ncclAllGather / peer copies instead of peer-memory mailboxes; read at
 *     pmt_comm_init and per pmt_merkle_tree_build_multi_dev call), PMT_EXCHANGE_TIMEOUT_MS (bounded waits inside kernels).
 *   - every host-buffer entry point that fails has drained its streams first: when it returns, the library no longer
 *     reads or writes the caller's buffers, whatever the status.
 */
#ifndef PMT_H
#define PMT_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMT_OK 0
#define PMT_E_INVALID_ARG (-1) /* null pointer, zero size where the reference would panic, width 0 ... */
#define PMT_E_NOT_POW2 (-2)    /* log2_strict panic: simple_merkle_tree.rs:30, [UPSTREAM] MerkleTree::new */
#define PMT_E_OOM (-3)
#define PMT_E_CUDA (-4)
#define PMT_E_NCCL (-6)  /* libnccl.so.2 not loadable, or an NCCL call failed */
#define PMT_E_RANGE (-5) /* index assert: simple_merkle_tree.rs:56,77; cap_height > log2 n; MMR leaf >= 2^30 (:264) */

typedef struct pmt_ctx pmt_ctx;

/* ---- context ---------------------------------------------------------------------------------------------------- */
int pmt_init(pmt_ctx** out, int device_id);
void pmt_destroy(pmt_ctx* ctx);
const char* pmt_last_error(const pmt_ctx* ctx);
const char* pmt_version(void);
int pmt_device_id(const pmt_ctx* ctx);
/* use an externally owned cudaStream_t (e.g. the caller's current stream); NULL restores the ctx's own stream.  Changing the
 * stream first waits for the work already enqueued on the old one (the ctx's device-side state assumes its launches run in order) */
int pmt_set_stream(pmt_ctx* ctx, void* cuda_stream);
void* pmt_get_stream(const pmt_ctx* ctx);
int pmt_sync(pmt_ctx* ctx);
/* number of CUDA kernels this ctx has launched since pmt_init (monotonic; bench.py reports the delta) */
uint64_t pmt_kernel_launches(const pmt_ctx* ctx);
/* optional per-launch timing: CUDA events on the launching stream around every kernel.  pmt_profile_read synchronises,
 * writes one text line per kernel ("name launches total_ms total_permutations") into buf, and resets the records. */
int pmt_profile_enable(pmt_ctx* ctx, int on);
int pmt_profile_read(pmt_ctx* ctx, char* buf, size_t cap);
/* device memory helpers for hosts without their own allocator (Rust shim); freed by pmt_free or pmt_destroy */
int pmt_malloc(pmt_ctx* ctx, size_t bytes, void** dptr_out);
int pmt_free(pmt_ctx* ctx, void* dptr);
int pmt_memcpy_h2d(pmt_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int pmt_memcpy_d2h(pmt_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);

/* Page-locked host memory for the host-buffer entry points.  Those pipeline H2D copies, hashing and D2H copies over chunks;
 * the overlap needs page-locked ("pinned") host buffers -- from pageable memory (a plain Rust Vec / malloc) the CUDA driver
 * stages every copy through its own bounce buffer and the pipeline serialises (measured: DESIGN.md 5).  Either lock
 * buffers the host already owns (pmt_host_register: cudaHostRegister; page-locking costs ~0.2 ms per MiB, so do it once
 * for buffers that are reused) or let the library allocate them (pmt_host_alloc: cudaHostAlloc, portable across devices;
 * in Rust the backing store of a Vec<T, A> with a custom allocator).  Results never depend on the kind of memory. */
int pmt_host_register(pmt_ctx* ctx, void* ptr, size_t bytes);
int pmt_host_unregister(pmt_ctx* ctx, void* ptr);
int pmt_host_alloc(pmt_ctx* ctx, size_t bytes, void** ptr_out);
int pmt_host_free(pmt_ctx* ctx, void* ptr);

/* ---- Hasher: [UPSTREAM] plonk/config.rs Hasher for PoseidonHash ------------------------------------------------------ */
/* n independent width-12 permutations (parity hook for Poseidon::poseidon); in/out n*12 felts */
int pmt_permute(pmt_ctx* ctx, const uint64_t* in, size_t n, uint64_t* out);
/* out[i] = PoseidonHash::two_to_one(l[i], r[i])   (simple_merkle_tree.rs:23,45,100,102; merkle_mountain_ranges.rs:111,238,240) */
int pmt_hash_two_to_one(pmt_ctx* ctx, const uint64_t* l, const uint64_t* r, size_t n, uint64_t* out);
/* out[i] = PoseidonHash::hash_or_noop(row i), rows row-major n_rows x width (simple_merkle_tree.rs:33,93;
 * merkle_mountain_ranges.rs:91,96,125,233,249): width <= 4 => canonicalise + zero pad, else overwrite-mode sponge */
int pmt_hash_or_noop(pmt_ctx* ctx, const uint64_t* rows, size_t n_rows, size_t width, uint64_t* out);
/* always the sponge (hash_n_to_hash_no_pad), also for width <= 4 */
int pmt_hash_no_pad(pmt_ctx* ctx, const uint64_t* rows, size_t n_rows, size_t width, uint64_t* out);
int pmt_permute_dev(pmt_ctx* ctx, const uint64_t* d_in, size_t n, uint64_t* d_out);
int pmt_hash_two_to_one_dev(pmt_ctx* ctx, const uint64_t* d_l, const uint64_t* d_r, size_t n, uint64_t* d_out);
/* noop_rule != 0: hash_or_noop, == 0: hash_no_pad */
int pmt_hash_rows_dev(pmt_ctx* ctx, const uint64_t* d_rows, size_t n_rows, size_t width, int noop_rule, uint64_t* d_out);

/* ---- simple tree: simple_merkle_tree.rs ------------------------------------------------------------------------------ */
/* MerkleTree::build (:28-51).  leaves: n felts, n a power of two >= 2.  levels_out: (2n-2) digests, level-major
 * (level 0 = n leaf digests, level 1 = n/2, ..., last = 2) = MerkleTree.tree flattened; root_out: 1 digest. */
int pmt_simple_tree_build(pmt_ctx* ctx, const uint64_t* leaves, size_t n, uint64_t* levels_out, uint64_t* root_out);
int pmt_simple_tree_build_dev(pmt_ctx* ctx, const uint64_t* d_leaves, size_t n, uint64_t* d_levels, uint64_t* d_root);
/* get_merkle_proof (:55-74) for a batch: siblings_out n_idx * log2(n) digests, bottom-up.  levels on the DEVICE.
 * The indices are on the device too, so the reference's assert!(leaf_index < n) (:56) cannot be a status code here: an
 * index >= n yields an all-zero proof (nothing outside `levels` is read).  The same holds for pmt_merkle_prove_dev;
 * pmt_mmr_prove_dev returns path_len 0 for it.  The host-buffer forms below return PMT_E_RANGE instead. */
int pmt_simple_tree_prove_dev(pmt_ctx* ctx, const uint64_t* d_levels, size_t n, const uint64_t* d_idx, size_t n_idx,
                              uint64_t* d_siblings_out);
/* verify_merkle_proof (:91-109) for a batch sharing one root: ok_out[i] = 1/0.  proofs: n_idx * path_len digests.
 * Exactly the reference's predicate: it folds by the parities of leaf_index >> i, i < path_len, and never looks at the
 * index bits above, so leaf_index + k 2^path_len verifies like leaf_index (upstream's verify_merkle_proof_to_cap below
 * rejects such an index: it selects cap[index >> path_len]). */
int pmt_simple_tree_verify_dev(pmt_ctx* ctx, const uint64_t* d_leaves, const uint64_t* d_idx, size_t n_idx,
                               const uint64_t* d_root, const uint64_t* d_proofs, size_t path_len, uint8_t* d_ok_out);

/* ---- plonky2 tree: [UPSTREAM] hash/merkle_tree.rs, hash/merkle_proofs.rs ------------------------------------------------ */
/* MerkleTree::new(leaves, cap_height).  leaves row-major n x width; digests_out 2(n - 2^h) digests in upstream's
 * interleaved layout (per cap subtree: left subtree || left digest || right digest || right subtree); cap_out 2^h. */
int pmt_merkle_tree_build(pmt_ctx* ctx, const uint64_t* leaves, size_t n, size_t width, uint32_t cap_height,
                          uint64_t* digests_out, uint64_t* cap_out);
int pmt_merkle_tree_build_dev(pmt_ctx* ctx, const uint64_t* d_leaves, size_t n, size_t width, uint32_t cap_height,
                              uint64_t* d_digests, uint64_t* d_cap);
/* MerkleTree::new over SEVERAL GPUs from one host process (the reference is a single process): n_ctx = 2^g
 * distinct contexts, normally one per device.  ctx r builds the subtree over leaves [r n/G, (r+1) n/G) on its device from
 * its own host thread, straight into its contiguous slice of digests_out; for cap_height < g the G subtree roots are
 * finished on ctxs[0] and the few digests above them placed between the slices.  Host buffers, synchronous, output
 * identical to pmt_merkle_tree_build.  (One process per GPU with an NCCL all_gather of the roots: DESIGN.md 7.) */
int pmt_merkle_tree_build_multi(pmt_ctx* const* ctxs, size_t n_ctx, const uint64_t* leaves, size_t n, size_t width,
                                uint32_t cap_height, uint64_t* digests_out, uint64_t* cap_out);
/* The same tree built straight from the prover's column-major LDE values: what [UPSTREAM] fri/oracle.rs
 * PolynomialBatch::from_values / from_coeffs feeds to MerkleTree::new inside circuit_data.prove
 * (/root/reference/src/mmr/mmr_plonky2_verifier.rs:148): leaves = reverse_index_bits(transpose(columns)), i.e.
 * leaf i = (col_0[rev(i)], ..., col_{w-1}[rev(i)]).  d_columns: w columns of n felts each, column-major.
 * bit_reverse != 0 applies the index reversal.  d_leaves_out (nullable): the row-major n x w leaves upstream keeps in
 * MerkleTree.leaves for openings.  The transpose never materialises unless d_leaves_out is given. */
int pmt_merkle_tree_build_from_columns_dev(pmt_ctx* ctx, const uint64_t* d_columns, size_t n, size_t width, int bit_reverse,
                                           uint32_t cap_height, uint64_t* d_leaves_out, uint64_t* d_digests, uint64_t* d_cap);
/* MerkleTree::prove for a batch: siblings_out n_idx * (log2 n - h) digests */
int pmt_merkle_prove_dev(pmt_ctx* ctx, const uint64_t* d_digests, size_t n, uint32_t cap_height, const uint64_t* d_idx,
                         size_t n_idx, uint64_t* d_siblings_out);
/* verify_merkle_proof_to_cap for a batch: leaves row-major n_idx x width */
int pmt_merkle_verify_dev(pmt_ctx* ctx, const uint64_t* d_leaf_rows, size_t width, const uint64_t* d_idx, size_t n_idx,
                          const uint64_t* d_cap, uint32_t cap_height, const uint64_t* d_proofs, size_t path_len,
                          uint8_t* d_ok_out);
/* multi-GPU finish: given the 2^g subtree roots gathered from the ranks (rank order), compute the g - h levels above
 * them.  d_top_out is level-major: 2^g/2, 2^g/4, ..., 2^h digests (2^g - 2^h in total); its last 2^h digests are the
 * cap.  cap_height == g: nothing to do (the roots are the cap). */
int pmt_top_levels_dev(pmt_ctx* ctx, const uint64_t* d_roots, size_t n_roots, uint32_t cap_height, uint64_t* d_top_out);

/* the same for `batch` independent sets of n_roots roots in one launch (sets n_roots digests apart in d_roots and
 * n_roots - 2^h digests apart in d_top_out): the rounds of a subtree-sharded MMR (one mountain per set bit of n) */
int pmt_top_levels_batch_dev(pmt_ctx* ctx, const uint64_t* d_roots, size_t batch, size_t n_roots, uint32_t cap_height,
                             uint64_t* d_top_out);

/* ---- subtree-sharded MerkleTree::new, device resident ----------------------------------------------------------
 * THE EXCHANGE.  Where every device can write its peers' memory (NVLink peer access inside one process, CUDA IPC mappings
 * between the processes of a communicator) the subtree roots travel through per-context MAILBOXES and the exchange is fused
 * with the levels above the roots in ONE kernel per context: it stores the context's root into every peer's mailbox with
 * plain NVLink stores + a flag, spins (bounded) on the flags in its own mailbox, and finishes the top levels.  No NCCL call, no
 * event, no copy on that path.  Otherwise -- PMT_EXCHANGE=nccl, no peer access, contexts that share a device, more than 64
 * ranks or 64 digests per rank -- the forms below fall back to what their comments describe (peer copies + events, or
 * ncclAllGather, then a separate finish launch); results are identical.  A peer that does not arrive within
 * PMT_EXCHANGE_TIMEOUT_MS (default 20 000) is reported by the next pmt_sync as PMT_E_NCCL instead of hanging the device.
 * (1) ONE PROCESS, several GPUs: n_ctx = 2^g distinct contexts, one per device.  ctx r builds the subtree over ITS leaves
 * (d_leaves[r]: n / n_ctx rows on ctx r's device) into d_digests[r] (its contiguous slice of upstream's `digests`:
 * 2 (n / n_ctx - max(1, 2^(h-g))) digests on its device).  cap_height >= g: ctx r's 2^(h-g) cap entries go straight to
 * d_cap + 4 r 2^(h-g) on ctxs[0]'s device.  Otherwise every ctx stores its subtree root into d_roots + 4 r on ctxs[0]'s
 * device -- a 32-byte peer-to-peer copy over NVLink on its own stream (cudaMemcpyPeerAsync; peer access is enabled by
 * the call when the devices allow it) -- ctxs[0]'s stream waits for the n_ctx copies (events, no host sync) and finishes
 * the g - h levels above the roots into d_top (level-major, n_ctx - 2^h digests; its last 2^h are the cap, also copied to
 * d_cap).  d_roots (n_ctx digests), d_top and d_cap (2^h digests) live on ctxs[0]'s device.  Only enqueues: the result is
 * complete after pmt_sync(ctxs[0]) (which orders after every other ctx's work through the events). */
int pmt_merkle_tree_build_multi_dev(pmt_ctx* const* ctxs, size_t n_ctx, const uint64_t* const* d_leaves, size_t n, size_t width,
                                    uint32_t cap_height, uint64_t* const* d_digests, uint64_t* d_roots, uint64_t* d_top,
                                    uint64_t* d_cap);
/* (2) ONE PROCESS PER GPU: the roots are exchanged with ncclAllGather on the ctx's own stream, inside the library (libnccl.so.2
 * is loaded with dlopen at pmt_comm_init: libpmt has no link-time NCCL dependency, and a process that already holds an
 * NCCL -- PyTorch's -- shares it).  pmt_nccl_unique_id: 128 bytes from ncclGetUniqueId, to be created on one rank and
 * handed to every rank by the host's own means; pmt_comm_init is collective over the `world` ranks (a power of two). */
int pmt_nccl_unique_id(pmt_ctx* ctx, void* id_out_128_bytes);
int pmt_comm_init(pmt_ctx* ctx, const void* unique_id_128_bytes, int rank, int world);
int pmt_comm_destroy(pmt_ctx* ctx);
/* 1 if the ranks of this ctx's communicator exchange through peer-memory mailboxes (decided unanimously in pmt_comm_init), 0 if
 * they use ncclAllGather */
int pmt_comm_uses_peer_memory(const pmt_ctx* ctx);
/* collective: this rank's n_total / world rows -> d_local_digests (its slice of `digests`).  cap_height >= log2 world: every
 * rank ends up with the whole cap in d_cap (2^h digests).  Otherwise d_roots (world digests) receives all subtree roots,
 * d_top (world - 2^h digests, level-major) the levels above them on EVERY rank, d_cap the cap.  One stream, no host sync:
 * local build -> ncclAllGather (32 bytes per rank over NVLink / NVSwitch) -> top levels. */
int pmt_merkle_tree_build_sharded_dev(pmt_ctx* ctx, const uint64_t* d_local_leaves, size_t n_total, size_t width,
                                      uint32_t cap_height, uint64_t* d_local_digests, uint64_t* d_roots, uint64_t* d_top,
                                      uint64_t* d_cap);

/* The subtree-sharded MMR with one process per GPU (pmt_comm_init first).  pmt_mmr_shard_plan is its pure index
 * math: every set bit 2^b >= world of n_total is one mountain, cut into `world` equal perfect sub-mountains of m[i] leaves
 * (round i, decreasing powers of two; at most 31 rounds); the bits below `world` are the tail of the last rank.  Rank r owns the
 * leaves [S_i + r m[i], S_i + (r + 1) m[i]) of every round (S_i = world (m[0] + .. + m[i-1])) and the last rank also the tail.
 * pmt_mmr_build_sharded_dev (collective, enqueues only): d_local_leaves = the rank's leaves in that order; outputs:
 * d_local_elements (pmt_mmr_size(sum m[i]) digests: sub-mountain i is the slice of the global `elements` starting at
 * pmt_mmr_size(S_i + r m[i])), d_tail_elements (last rank, pmt_mmr_size(tail) digests), d_gathered (world x slots digests,
 * slots = rounds + popcount(tail): every rank's sub-mountain roots and the tail's peaks), d_tops (rounds x (world - 1) digests:
 * the levels above each round's roots, level-major; global positions as for pmt_top_levels_dev) and d_peaks (slots digests:
 * get_peaks() of the whole MMR, replicated).  One batch append, one ncclAllGather, one batched finish, one stream. */
int pmt_mmr_shard_plan(size_t n_total, size_t world, uint32_t* n_rounds, size_t* m_out_64, size_t* tail);
int pmt_mmr_build_sharded_dev(pmt_ctx* ctx, const uint64_t* d_local_leaves, size_t n_total, uint64_t* d_local_elements,
                              uint64_t* d_tail_elements, uint64_t* d_gathered, uint64_t* d_tops, uint64_t* d_peaks);

/* ---- MMR: merkle_mountain_ranges.rs ------------------------------------------------------------------------------------ */
/* number of elements of an MMR with n leaves = 2n - popcount(n) */
size_t pmt_mmr_size(size_t n_leaves);
/* get_mmr_index (:257-270) = 2i - popcount(i) */
size_t pmt_mmr_index(size_t leaf_normal_index);
/* batch of MMR::add_leaf (:89-120): append m single-felt leaves to an MMR that already has n_before leaves.
 * elements: post-order array with capacity >= pmt_mmr_size(n_before + m) digests, first pmt_mmr_size(n_before) valid.
 * The host-buffer form reads only the popcount(n_before) old peaks of `elements` and writes only the new elements; the
 * device holds just those (O(m + log n_before) memory and PCIe traffic, like the O(log n) add_leaf it batches).  It is
 * one synchronous round trip per call: feed leaves in batches, not one by one. */
int pmt_mmr_extend(pmt_ctx* ctx, uint64_t* elements, size_t n_before, const uint64_t* new_leaves, size_t m);
int pmt_mmr_extend_dev(pmt_ctx* ctx, uint64_t* d_elements, size_t n_before, const uint64_t* d_new_leaves, size_t m);
/* The same batch append over SEVERAL GPUs from one host process: n_ctx distinct contexts (any count), normally one per
 * device.  The appended leaves are cut into aligned blocks of 2^b leaves -- perfect sub-mountains, each one contiguous
 * slice of `elements` -- which context j mod n_ctx builds from its own host thread; the nodes above the block roots are
 * finished on ctxs[0].  Host buffers, synchronous, output identical to pmt_mmr_extend (which it calls when there is one
 * context or fewer than 4096 leaves per context).  pmt_mmr_multi_plan is the pure index math of the cut (no device;
 * returns 1 and b, the first and the last block boundary if the append is split, else 0). */
int pmt_mmr_extend_multi(pmt_ctx* const* ctxs, size_t n_ctx, uint64_t* elements, size_t n_before, const uint64_t* new_leaves,
                         size_t m);
int pmt_mmr_multi_plan(size_t n_before, size_t m, size_t n_ctx, uint32_t* log2_block, size_t* first_aligned,
                       size_t* last_aligned);
/* get_peaks (:179-200): peaks_out up to 64 digests, largest mountain first; *n_peaks_out = popcount(n_leaves) */
int pmt_mmr_peaks_dev(pmt_ctx* ctx, const uint64_t* d_elements, size_t n_leaves, uint64_t* d_peaks_out,
                      uint32_t* n_peaks_out);
/* bagging_the_peaks (:122-127): hash_or_noop over the flattened peaks */
int pmt_mmr_bag_dev(pmt_ctx* ctx, const uint64_t* d_elements, size_t n_leaves, uint64_t* d_root_out);
/* get_proof_normal_index (:203-223) for a batch of NORMAL leaf indices.  Per proof: up to 32 (sibling, on_left) entries
 * at stride 32 digests / 32 bytes; path_len_out[i] = height of the leaf's mountain.  Peaks: use pmt_mmr_peaks_dev. */
int pmt_mmr_prove_dev(pmt_ctx* ctx, const uint64_t* d_elements, size_t n_leaves, const uint64_t* d_leaf_idx,
                      size_t n_idx, uint64_t* d_siblings_out, uint8_t* d_on_left_out, uint32_t* d_path_len_out);
/* MMR_proof::verify (:232-252) for a batch sharing peaks + root: status_out[i] = 1 true, 0 false,
 * -1 = the reference would panic (subtree root not among the peaks, assert! at :245).  A proof is untrusted input:
 * path_len[i] > 32 (more than the layout holds; an MMR has < 2^30 leaves) is -1 as well and nothing is folded. */
int pmt_mmr_verify_dev(pmt_ctx* ctx, const uint64_t* d_leaves, size_t n_idx, const uint64_t* d_siblings,
                       const uint8_t* d_on_left, const uint32_t* d_path_len, const uint64_t* d_peaks, uint32_t n_peaks,
                       const uint64_t* d_root, int8_t* d_status_out);
/* host-buffer forms: only the <= 32 peak digests are touched (peaks: a gather on the host, nothing to compute; bag: the
 * peaks are uploaded and hashed on the GPU) */
int pmt_mmr_bag(pmt_ctx* ctx, const uint64_t* elements, size_t n_leaves, uint64_t* root_out);
int pmt_mmr_peaks(pmt_ctx* ctx, const uint64_t* elements, size_t n_leaves, uint64_t* peaks_out, uint32_t* n_peaks_out);

/* ---- host-buffer proofs and verification ---------------------------------------------------------------------------------
 * Proofs are gathers from the arrays the host-buffer builders filled: done on the host, no GPU work (the reference does
 * the same lookups after cloning the whole tree: simple_merkle_tree.rs:55-74, merkle_mountain_ranges.rs:147-223).  An
 * index the reference would panic on is PMT_E_RANGE.  Output layouts as for the *_dev forms. */
int pmt_simple_tree_prove(pmt_ctx* ctx, const uint64_t* levels, size_t n, const uint64_t* idx, size_t n_idx,
                          uint64_t* siblings_out);
int pmt_merkle_prove(pmt_ctx* ctx, const uint64_t* digests, size_t n, uint32_t cap_height, const uint64_t* idx, size_t n_idx,
                     uint64_t* siblings_out);
int pmt_mmr_prove(pmt_ctx* ctx, const uint64_t* elements, size_t n_leaves, const uint64_t* leaf_idx, size_t n_idx,
                  uint64_t* siblings_out, uint8_t* on_left_out, uint32_t* path_len_out);
/* verify_merkle_proof (:91-109), verify_merkle_proof_to_cap, MMR_proof::verify (:232-252) for a batch in HOST buffers: the
 * batch is uploaded, every path is folded on the GPU, the verdicts come back; synchronous. */
int pmt_simple_tree_verify(pmt_ctx* ctx, const uint64_t* leaves, const uint64_t* idx, size_t n_idx, const uint64_t* root,
                           const uint64_t* proofs, size_t path_len, uint8_t* ok_out);
int pmt_merkle_verify(pmt_ctx* ctx, const uint64_t* leaf_rows, size_t width, const uint64_t* idx, size_t n_idx,
                      const uint64_t* cap, uint32_t cap_height, const uint64_t* proofs, size_t path_len, uint8_t* ok_out);
int pmt_mmr_verify(pmt_ctx* ctx, const uint64_t* leaves, size_t n_idx, const uint64_t* siblings, const uint8_t* on_left,
                   const uint32_t* path_len, const uint64_t* peaks, uint32_t n_peaks, const uint64_t* root,
                   int8_t* status_out);

#ifdef __cplusplus
}
#endif
#endif /* PMT_H */
