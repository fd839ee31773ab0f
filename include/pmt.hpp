// pmt.hpp -- C++17 host mirror of the reference's Rust interface over the C ABI of libpmt (include/pmt.h).
//
// The reference is compiled code (Rust) and this image has no Rust toolchain, so the host side a Rust maintainer would
// write exists here in C++: the same type names, field names, argument meaning and error
// behaviour as
//     /root/reference/src/simple_merkle_tree/simple_merkle_tree.rs   MerkleTree{count_levels, tree, root} :11-16,
//                                                                    build :28, get_merkle_proof :55,
//                                                                    get_in_between_hashes :76, verify_merkle_proof :91
//     /root/reference/src/mmr/merkle_mountain_ranges.rs              MMR{elements} :8-12, MMR_proof :15-23,
//                                                                    get_heights_bitmap_for_mmr_size :39, add_leaf :89,
//                                                                    bagging_the_peaks :122, get_peaks :179,
//                                                                    get_proof_normal_index :203, get_proof :209,
//                                                                    MMR_proof::verify :232, get_mmr_index :257
//     [UPSTREAM plonky2 hash/merkle_tree.rs, hash/merkle_proofs.rs]  MerkleTree{leaves, digests, cap}, MerkleCap,
//                                                                    MerkleProof{siblings}, new, prove,
//                                                                    verify_merkle_proof_to_cap
// State lives where the reference keeps it (host vectors with the reference's layouts); every hash is computed by the
// GPU engine through the C ABI, and proofs are gathers from the host arrays exactly as in the reference.  A panic of the
// reference (log2_strict, index asserts, the assert! in MMR_proof::verify) is a pmt::Error here.  Header-only; link
// with -lpmt.  There is no CPU fallback: Engine's constructor throws when libpmt finds no CUDA device.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <exception>
#include <stdexcept>
#include <memory>
#include <string>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

#include "pmt.h"

namespace pmt {

using F = uint64_t;  // GoldilocksField(u64): any u64 on input, canonical on output

struct HashOut {  // plonky2 HashOut<GoldilocksField>.  Trivial: `HashOut h{}` is zero, `HashOut h;` is filled by the next call
  std::array<F, 4> elements;
  bool operator==(const HashOut& o) const { return elements == o.elements; }
  bool operator!=(const HashOut& o) const { return !(*this == o); }
};
static_assert(sizeof(HashOut) == 32, "a digest is 4 consecutive u64 in every libpmt buffer");

// Vec<HashOut> for the big outputs (1 GiB of digests for 2^24 leaves).  std::vector::resize would write zeros into every
// element libpmt is about to overwrite -- a serial pass over fresh pages that costs more than the GPU build.  This allocator
// default-initialises instead (the C++ stand-in for Rust's Vec::with_capacity + set_len after the FFI call has filled it).
template <class T>
struct NoInitAlloc : std::allocator<T> {
  template <class U> struct rebind { using other = NoInitAlloc<U>; };
  NoInitAlloc() = default;
  template <class U> NoInitAlloc(const NoInitAlloc<U>&) {}
  template <class U> void construct(U* p) noexcept { ::new (static_cast<void*>(p)) U; }
  template <class U, class... A> void construct(U* p, A&&... a) { ::new (static_cast<void*>(p)) U(std::forward<A>(a)...); }
};
using Digests = std::vector<HashOut, NoInitAlloc<HashOut>>;
template <class A, class B>
typename std::enable_if<!std::is_same<A, B>::value, bool>::type operator==(const std::vector<HashOut, A>& x, const std::vector<HashOut, B>& y) {
  return x.size() == y.size() && std::equal(x.begin(), x.end(), y.begin());
}

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

// One pmt_ctx (one CUDA device, one stream).  Not thread-safe, like the ctx.
class Engine {
 public:
  explicit Engine(int device_id = 0) {
    const int rc = pmt_init(&ctx_, device_id);
    if (rc != PMT_OK) throw Error(rc, "pmt_init failed: no usable CUDA device (libpmt has no CPU fallback)");
  }
  ~Engine() { pmt_destroy(ctx_); }
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;
  pmt_ctx* ctx() const { return ctx_; }
  void check(int rc) const {
    if (rc != PMT_OK) throw Error(rc, std::string("libpmt error ") + std::to_string(rc) + ": " + pmt_last_error(ctx_));
  }
  uint64_t kernel_launches() const { return pmt_kernel_launches(ctx_); }

  // ---- Hasher for PoseidonHash ([UPSTREAM] plonk/config.rs) ----
  HashOut two_to_one(const HashOut& l, const HashOut& r) const {
    HashOut out;
    check(pmt_hash_two_to_one(ctx_, l.elements.data(), r.elements.data(), 1, out.elements.data()));
    return out;
  }
  HashOut hash_or_noop(const std::vector<F>& in) const {
    HashOut out;
    if (in.empty()) return out;  // hash_or_noop(&[]) pads nothing: the zero digest
    check(pmt_hash_or_noop(ctx_, in.data(), 1, in.size(), out.elements.data()));
    return out;
  }
  HashOut hash_no_pad(const std::vector<F>& in) const {
    HashOut out;
    check(pmt_hash_no_pad(ctx_, in.data(), 1, in.size(), out.elements.data()));
    return out;
  }

 private:
  pmt_ctx* ctx_ = nullptr;
};

namespace detail {
template <class A> inline const uint64_t* words(const std::vector<HashOut, A>& v) { return reinterpret_cast<const uint64_t*>(v.data()); }
template <class A> inline uint64_t* words(std::vector<HashOut, A>& v) { return reinterpret_cast<uint64_t*>(v.data()); }
// Vec<Vec<F>> -> one row-major array, by a few threads (16 M four-felt rows scattered over the heap: 170 ms on one thread)
inline std::vector<F, NoInitAlloc<F>> flatten(const std::vector<std::vector<F>>& rows, size_t w) {
  const size_t n = rows.size();
  std::vector<F, NoInitAlloc<F>> flat(n * w);
  auto part = [&](size_t a, size_t b) {
    for (size_t i = a; i < b; i++) {
      if (rows[i].size() != w) throw Error(PMT_E_INVALID_ARG, "MerkleTree::new: ragged leaves");
      std::copy(rows[i].begin(), rows[i].end(), flat.begin() + i * w);
    }
  };
  unsigned t = n * w >= (size_t(1) << 20) ? std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2)) : 1;
  if (t <= 1) { part(0, n); return flat; }
  std::vector<std::thread> pool;
  std::vector<std::exception_ptr> err(t);
  for (unsigned k = 0; k < t; k++)
    pool.emplace_back([&, k] { try { part(n * k / t, n * (k + 1) / t); } catch (...) { err[k] = std::current_exception(); } });
  for (auto& th : pool) th.join();
  for (auto& e : err) if (e) std::rethrow_exception(e);
  return flat;
}
inline unsigned popcount(uint64_t x) { return (unsigned)__builtin_popcountll(x); }
inline bool is_pow2(size_t n) { return n && !(n & (n - 1)); }
inline unsigned log2_strict(size_t n) {
  if (!is_pow2(n)) throw Error(PMT_E_NOT_POW2, "log2_strict: " + std::to_string(n) + " is not a power of two");
  return (unsigned)__builtin_ctzll(n);
}
}  // namespace detail

// ====================================================================================================================
namespace simple_merkle_tree {

struct MerkleTree {
  size_t count_levels = 0;
  std::vector<std::vector<HashOut>> tree;  // level 0 = leaf digests ... level count_levels-1 = 2 digests
  HashOut root;

  // simple_merkle_tree.rs:28-51
  static MerkleTree build(const Engine& e, const std::vector<F>& leaves) {
    const size_t n = leaves.size();
    MerkleTree t;
    t.count_levels = detail::log2_strict(n);  // :30
    if (n < 2) throw Error(PMT_E_INVALID_ARG, "MerkleTree::build needs at least 2 leaves (simple_merkle_tree.rs:38)");
    std::vector<HashOut> levels(2 * n - 2);
    e.check(pmt_simple_tree_build(e.ctx(), leaves.data(), n, detail::words(levels), t.root.elements.data()));
    size_t off = 0;
    for (size_t m = n; m >= 2; off += m, m /= 2) t.tree.emplace_back(levels.begin() + off, levels.begin() + off + m);
    return t;
  }

  // :55-74 -- siblings bottom-up
  std::vector<HashOut> get_merkle_proof(size_t leaf_index) const {
    if (leaf_index >= tree[0].size()) throw Error(PMT_E_RANGE, "assert!(leaf_index < self.tree[0].len()) (simple_merkle_tree.rs:56)");
    std::vector<HashOut> proof;
    size_t idx = leaf_index;
    for (size_t i = 0; i < count_levels; i++, idx >>= 1) proof.push_back(tree[i][idx ^ 1]);
    return proof;
  }

  // :76-87 -- the nodes on the path, then the root
  std::vector<HashOut> get_in_between_hashes(size_t leaf_index) const {
    if (leaf_index >= tree[0].size()) throw Error(PMT_E_RANGE, "assert!(leaf_index < self.tree[0].len()) (simple_merkle_tree.rs:77)");
    std::vector<HashOut> out;
    size_t idx = leaf_index >> 1;
    for (size_t i = 1; i < count_levels; i++, idx >>= 1) out.push_back(tree[i][idx]);
    out.push_back(root);
    return out;
  }
};

// batch of verify_merkle_proof (:91-109) against one root: one kernel launch folds every path
inline std::vector<bool> verify_merkle_proofs(const Engine& e, const std::vector<F>& leaves, const std::vector<uint64_t>& leaf_indices,
                                              const HashOut& root, const std::vector<std::vector<HashOut>>& proofs) {
  const size_t n = leaves.size();
  if (leaf_indices.size() != n || proofs.size() != n) throw Error(PMT_E_INVALID_ARG, "verify_merkle_proofs: size mismatch");
  if (n == 0) return {};
  const size_t path_len = proofs[0].size();
  std::vector<HashOut> flat;
  for (const auto& p : proofs) {
    if (p.size() != path_len) throw Error(PMT_E_INVALID_ARG, "verify_merkle_proofs: proofs of different lengths");
    flat.insert(flat.end(), p.begin(), p.end());
  }
  std::vector<uint8_t> ok(n);
  e.check(pmt_simple_tree_verify(e.ctx(), leaves.data(), leaf_indices.data(), n, root.elements.data(), detail::words(flat), path_len, ok.data()));
  return std::vector<bool>(ok.begin(), ok.end());
}

inline bool verify_merkle_proof(const Engine& e, F leaf, size_t leaf_index, const HashOut& root, const std::vector<HashOut>& hashes) {
  return verify_merkle_proofs(e, {leaf}, {(uint64_t)leaf_index}, root, {hashes})[0];
}

}  // namespace simple_merkle_tree

// ====================================================================================================================
namespace mmr {

// merkle_mountain_ranges.rs:39-81 -> (bitmap of mountain heights, elements that do not form a full mountain)
inline std::pair<uint64_t, size_t> get_heights_bitmap_for_mmr_size(size_t mmr_size) {
  if (mmr_size == 0) return {0, 0};
  size_t subtree = (~size_t(0)) >> __builtin_clzll(mmr_size);  // 2^(bit length) - 1
  uint64_t peaks = 0;
  size_t left = mmr_size;
  for (; subtree > 0; subtree >>= 1) {
    peaks <<= 1;
    if (left >= subtree) { peaks |= 1; left -= subtree; }
  }
  return {peaks, left};
}

// :257-270 (= 2 i - popcount(i)); the reference computes in i32, so i must stay below 2^30 (:264)
inline size_t get_mmr_index(size_t leaf_normal_index) {
  if (leaf_normal_index >= (size_t(1) << 30)) throw Error(PMT_E_RANGE, "get_mmr_index: i32 overflow in the reference (merkle_mountain_ranges.rs:264)");
  return pmt_mmr_index(leaf_normal_index);
}

struct MMR_proof {
  size_t mmr_size = 0;                                  // elements.len() when the proof was made
  std::vector<std::pair<HashOut, bool>> merkle_proof;   // (sibling, sibling_on_left), bottom-up
  std::vector<HashOut> peaks;

  // :232-252.  Throws where the reference's assert!(self.peaks.contains(&next_hash)) (:245) panics.
  bool verify(const Engine& e, F leaf, const HashOut& root) const {
    if (merkle_proof.size() > 32) throw Error(PMT_E_RANGE, "MMR_proof: path longer than 32");
    std::array<HashOut, 32> sib{};
    std::array<uint8_t, 32> left{};
    for (size_t j = 0; j < merkle_proof.size(); j++) { sib[j] = merkle_proof[j].first; left[j] = merkle_proof[j].second; }
    const uint32_t len = (uint32_t)merkle_proof.size();
    int8_t status = 0;
    e.check(pmt_mmr_verify(e.ctx(), &leaf, 1, reinterpret_cast<const uint64_t*>(sib.data()), left.data(), &len, detail::words(peaks),
                           (uint32_t)peaks.size(), root.elements.data(), &status));
    if (status < 0) throw Error(PMT_E_INVALID_ARG, "assert!(self.peaks.contains(&next_hash)) (merkle_mountain_ranges.rs:245)");
    return status == 1;
  }
};

struct MMR {
  std::vector<HashOut> elements;  // post-order, :8-12
  size_t n_leaves = 0;            // not in the reference (it re-derives it from elements.len()); kept to avoid the search

  static MMR new_() { return MMR(); }  // :84

  // batch of add_leaf (:89-120): one engine call hashes every node the new leaves complete
  void extend(const Engine& e, const std::vector<F>& leaves) {
    if (leaves.empty()) return;
    if (n_leaves + leaves.size() > (size_t(1) << 30)) throw Error(PMT_E_RANGE, "MMR leaf count > 2^30 (merkle_mountain_ranges.rs:264)");
    elements.resize(pmt_mmr_size(n_leaves + leaves.size()));
    e.check(pmt_mmr_extend(e.ctx(), detail::words(elements), n_leaves, leaves.data(), leaves.size()));
    n_leaves += leaves.size();
  }
  // the same batch over several GPUs from this one process (pmt_mmr_extend_multi): distinct engines, normally one per device
  void extend_multi(const std::vector<const Engine*>& engines, const std::vector<F>& leaves) {
    if (engines.empty()) throw Error(PMT_E_INVALID_ARG, "MMR::extend_multi: no engine");
    if (leaves.empty()) return;
    if (n_leaves + leaves.size() > (size_t(1) << 30)) throw Error(PMT_E_RANGE, "MMR leaf count > 2^30 (merkle_mountain_ranges.rs:264)");
    std::vector<pmt_ctx*> ctxs;
    for (const Engine* e : engines) ctxs.push_back(e ? e->ctx() : nullptr);
    elements.resize(pmt_mmr_size(n_leaves + leaves.size()));
    engines[0]->check(pmt_mmr_extend_multi(ctxs.data(), ctxs.size(), detail::words(elements), n_leaves, leaves.data(), leaves.size()));
    n_leaves += leaves.size();
  }
  void add_leaf(const Engine& e, F leaf) { extend(e, {leaf}); }

  // :179-200 -- one peak per set bit of the leaf count, largest mountain first
  std::vector<HashOut> get_peaks() const {
    std::vector<HashOut> peaks;
    size_t pos = 0;
    for (int b = 63; b >= 0; b--)
      if (n_leaves >> b & 1) { pos += (size_t(2) << b) - 1; peaks.push_back(elements[pos - 1]); }
    return peaks;
  }

  // :122-127 -- hash_or_noop over the flattened peaks
  HashOut bagging_the_peaks(const Engine& e) const {
    HashOut root;
    if (n_leaves == 0) return root;
    e.check(pmt_mmr_bag(e.ctx(), detail::words(elements), n_leaves, root.elements.data()));
    return root;
  }

  // :203-205
  MMR_proof get_proof_normal_index(size_t normal_index) const {
    if (normal_index >= n_leaves) throw Error(PMT_E_RANGE, "MMR::get_proof: leaf index out of range");
    // the mountain holding the leaf: mountains are the set bits of n_leaves, high to low
    size_t first_leaf = 0;
    unsigned height = 0;
    for (int b = 63; b >= 0; b--)
      if (n_leaves >> b & 1) {
        if (normal_index < first_leaf + (size_t(1) << b)) { height = (unsigned)b; break; }
        first_leaf += size_t(1) << b;
      }
    MMR_proof proof;
    proof.mmr_size = elements.size();
    const size_t local = normal_index - first_leaf;
    for (unsigned l = 0; l < height; l++) {
      const size_t k = (normal_index >> l) ^ 1;                 // sibling at height l, index among the height-l nodes
      const size_t last = ((k + 1) << l) - 1;                   // its last leaf
      const size_t pos = 2 * last - detail::popcount(last) + l;  // post-order position
      proof.merkle_proof.emplace_back(elements[pos], (local >> l & 1) != 0);
    }
    proof.peaks = get_peaks();
    return proof;
  }

  // :209-223 -- mmr_index must be the position of a leaf
  MMR_proof get_proof(size_t mmr_index) const {
    size_t lo = 0, hi = mmr_index + 1;  // invert 2 i - popcount(i), strictly increasing in i
    while (lo < hi) {
      const size_t mid = (lo + hi) / 2;
      if (2 * mid - detail::popcount(mid) < mmr_index) lo = mid + 1; else hi = mid;
    }
    if (2 * lo - detail::popcount(lo) != mmr_index) throw Error(PMT_E_INVALID_ARG, "MMR::get_proof: mmr_index is not a leaf position");
    return get_proof_normal_index(lo);
  }
};

}  // namespace mmr

// ====================================================================================================================
namespace plonky2 {

struct MerkleCap { std::vector<HashOut> hashes; size_t height() const { return detail::log2_strict(hashes.size()); } };
struct MerkleProof { std::vector<HashOut> siblings; };

struct MerkleTree {
  std::vector<std::vector<F>> leaves;
  Digests digests;  // upstream's layout: per cap subtree, left subtree || left digest || right digest || right subtree
  MerkleCap cap;

  // [UPSTREAM] MerkleTree::new(leaves, cap_height)
  static MerkleTree new_(const Engine& e, std::vector<std::vector<F>> leaves, size_t cap_height) {
    const size_t n = leaves.size();
    const unsigned log2n = detail::log2_strict(n);
    if (cap_height > log2n) throw Error(PMT_E_RANGE, "cap_height " + std::to_string(cap_height) + " > log2(leaves.len())");
    const size_t w = leaves[0].size();
    const auto flat = detail::flatten(leaves, w);
    MerkleTree t;
    t.digests.resize(2 * (n - (size_t(1) << cap_height)));
    t.cap.hashes.resize(size_t(1) << cap_height);
    uint64_t dummy[4];
    e.check(pmt_merkle_tree_build(e.ctx(), flat.data(), n, w, (uint32_t)cap_height, t.digests.empty() ? dummy : detail::words(t.digests),
                                  detail::words(t.cap.hashes)));
    t.leaves = std::move(leaves);
    return t;
  }

  // The same constructor over several GPUs from this one process (pmt_merkle_tree_build_multi): a power-of-two list of
  // distinct engines, normally one per device; engine r builds the subtree over leaves [r n/G, (r+1) n/G) on its own device.
  static MerkleTree new_multi(const std::vector<const Engine*>& engines, std::vector<std::vector<F>> leaves, size_t cap_height) {
    if (engines.empty()) throw Error(PMT_E_INVALID_ARG, "MerkleTree::new_multi: no engine");
    const size_t n = leaves.size();
    const unsigned log2n = detail::log2_strict(n);
    if (cap_height > log2n) throw Error(PMT_E_RANGE, "cap_height " + std::to_string(cap_height) + " > log2(leaves.len())");
    const size_t w = leaves[0].size();
    const auto flat = detail::flatten(leaves, w);
    std::vector<pmt_ctx*> ctxs;
    for (const Engine* e : engines) ctxs.push_back(e ? e->ctx() : nullptr);
    MerkleTree t;
    t.digests.resize(2 * (n - (size_t(1) << cap_height)));
    t.cap.hashes.resize(size_t(1) << cap_height);
    uint64_t dummy[4];
    engines[0]->check(pmt_merkle_tree_build_multi(ctxs.data(), ctxs.size(), flat.data(), n, w, (uint32_t)cap_height,
                                                  t.digests.empty() ? dummy : detail::words(t.digests), detail::words(t.cap.hashes)));
    t.leaves = std::move(leaves);
    return t;
  }

  // [UPSTREAM] MerkleTree::prove: siblings bottom-up inside the leaf's cap subtree
  MerkleProof prove(size_t leaf_index) const {
    const size_t n = leaves.size();
    if (leaf_index >= n) throw Error(PMT_E_RANGE, "MerkleTree::prove: leaf index out of range");
    const unsigned levels = detail::log2_strict(n) - (unsigned)cap.height();  // L = height of one cap subtree
    const size_t subtree_len = (size_t(2) << levels) - 2, c = leaf_index >> levels, local = leaf_index & ((size_t(1) << levels) - 1);
    MerkleProof p;
    for (unsigned l = 0; l < levels; l++) {
      const size_t k = (local >> l) ^ 1;
      p.siblings.push_back(digests[c * subtree_len + 2 * (((k >> 1) << (l + 1)) + (size_t(1) << l) - 1) + (k & 1)]);
    }
    return p;
  }
};

// [UPSTREAM] hash/merkle_proofs.rs verify_merkle_proof_to_cap; false where upstream returns Err
inline bool verify_merkle_proof_to_cap(const Engine& e, const std::vector<F>& leaf_data, size_t leaf_index, const MerkleCap& cap,
                                       const MerkleProof& proof) {
  const uint64_t idx = leaf_index;
  uint8_t ok = 0;
  e.check(pmt_merkle_verify(e.ctx(), leaf_data.data(), leaf_data.size(), &idx, 1, detail::words(cap.hashes), (uint32_t)cap.height(),
                            detail::words(proof.siblings), proof.siblings.size(), &ok));
  return ok == 1;
}

}  // namespace plonky2
}  // namespace pmt
