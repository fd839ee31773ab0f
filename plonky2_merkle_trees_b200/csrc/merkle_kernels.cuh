// merkle_kernels.cuh -- tree / MMR kernels over the per-thread Poseidon permutation (sm_100a).
//
// HBM layouts (all digests are 4 x u64 = 32 B, canonical):
//   LevelMajor  simple tree, /root/reference/src/simple_merkle_tree/simple_merkle_tree.rs:12-16 (MerkleTree.tree
//               flattened): level l (0 = leaf digests) starts at digest 2n - (2n >> l); the root is a separate output.
//   Plonky2     [UPSTREAM hash/merkle_tree.rs] `digests`: per cap subtree  left subtree || left digest || right digest ||
//               right subtree;  node (l, k') of a subtree sits at 2*(((k'>>1) << (l+1)) + (1<<l) - 1) + (k'&1);
//               subtree roots go to `cap`.
//   Mmr         /root/reference/src/mmr/merkle_mountain_ranges.rs:8-12 `elements`, post-order: node (l, k) covering leaves
//               [k 2^l, (k+1) 2^l) sits at 2*last - popcount(last) + l, last = (k+1) 2^l - 1; children at pos-2^l, pos-1.
// In every layout the two children of a node are read with 128-bit loads and the parent is written once; nodes are
// never re-read by the level that produced them, so HBM traffic is the algorithmic 96 B per two_to_one.
#pragma once
#include "poseidon.cuh"
#include "poseidon_coop.cuh"

namespace pmt {

using poseidon::WIDTH;
using poseidon::permute;   // one state per thread (poseidon.cuh); the four-threads-per-state form is poseidon::coop

struct Digest { uint64_t v[4]; };

__device__ __forceinline__ Digest load_digest(const uint64_t* __restrict__ p) {
  const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(p);
  const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(p + 2);
  Digest d; d.v[0] = a.x; d.v[1] = a.y; d.v[2] = b.x; d.v[3] = b.y;
  return d;
}
__device__ __forceinline__ void store_digest(uint64_t* __restrict__ p, const Digest& d) {
  *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2(d.v[0], d.v[1]);
  *reinterpret_cast<ulonglong2*>(p + 2) = make_ulonglong2(d.v[2], d.v[3]);
}

// [UPSTREAM hash/hashing.rs compress]: perm(l || r || 0^4)[0..4)
__device__ __forceinline__ Digest two_to_one(const Digest& l, const Digest& r) {
  uint64_t s[WIDTH] = {l.v[0], l.v[1], l.v[2], l.v[3], r.v[0], r.v[1], r.v[2], r.v[3], 0, 0, 0, 0};
  permute(s);
  Digest d;
#pragma unroll
  for (int i = 0; i < 4; i++) d.v[i] = gl::canonical(s[i]);
  return d;
}

// [UPSTREAM hash/hashing.rs hash_n_to_m_no_pad]: overwrite-mode sponge, rate 8, no padding; row = w felts at `row`
__device__ __forceinline__ Digest hash_no_pad(const uint64_t* __restrict__ row, size_t w) {
  uint64_t s[WIDTH];
#pragma unroll
  for (int i = 0; i < WIDTH; i++) s[i] = 0;
  for (size_t off = 0; off < w; off += 8) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      if (off + i < w) s[i] = __ldg(row + off + i);
    permute(s);
  }
  Digest d;
#pragma unroll
  for (int i = 0; i < 4; i++) d.v[i] = gl::canonical(s[i]);
  return d;
}

// [UPSTREAM plonk/config.rs Hasher::hash_or_noop]: <= 4 felts => identity with zero padding (canonicalised)
__device__ __forceinline__ Digest hash_or_noop(const uint64_t* __restrict__ row, size_t w) {
  if (w <= 4) {
    Digest d;
#pragma unroll
    for (int i = 0; i < 4; i++) d.v[i] = (size_t)i < w ? gl::canonical(__ldg(row + i)) : 0ull;
    return d;
  }
  return hash_no_pad(row, w);
}

// ---------------------------------------------------------------------------------------------------------------
// layouts
// ---------------------------------------------------------------------------------------------------------------
struct LevelMajor {
  uint64_t* base;   // (2n - 2) digests
  uint64_t* root;   // 1 digest
  size_t n;         // leaves
  int top;          // log2 n  (level `top` = root)
  __device__ __forceinline__ uint64_t* at(int l, size_t k) const {
    if (l == top) return root;
    return base + 4 * ((2 * n - ((2 * n) >> l)) + k);
  }
  __device__ __forceinline__ void children(int l, size_t k, const uint64_t*& a, const uint64_t*& b) const {
    a = at(l - 1, 2 * k); b = a + 4;
  }
  __device__ __forceinline__ LevelMajor for_set(unsigned) const { return *this; }
};

struct Plonky2 {
  uint64_t* digests;  // 2 (n - 2^h) digests
  uint64_t* cap;      // 2^h digests
  int sub_levels;     // L = log2 n - h  (level L = subtree roots -> cap)
  __device__ __forceinline__ uint64_t* at(int l, size_t k) const {
    if (l == sub_levels) return cap + 4 * k;
    const int per = sub_levels - l;                       // log2(nodes of this level per subtree)
    const size_t c = k >> per, kk = k & (((size_t)1 << per) - 1);
    const size_t sub_len = ((size_t)2 << sub_levels) - 2;
    const size_t idx = 2 * (((kk >> 1) << (l + 1)) + ((size_t)1 << l) - 1) + (kk & 1);
    return digests + 4 * (c * sub_len + idx);
  }
  __device__ __forceinline__ void children(int l, size_t k, const uint64_t*& a, const uint64_t*& b) const {
    a = at(l - 1, 2 * k); b = a + 4;                      // siblings are adjacent in this layout
  }
  __device__ __forceinline__ Plonky2 for_set(unsigned) const { return *this; }
};

struct Mmr {
  uint64_t* elements;
  __device__ __forceinline__ static size_t pos(int l, size_t k) {
    const size_t last = ((k + 1) << l) - 1;
    return 2 * last - (size_t)__popcll((unsigned long long)last) + (size_t)l;
  }
  __device__ __forceinline__ uint64_t* at(int l, size_t k) const { return elements + 4 * pos(l, k); }
  __device__ __forceinline__ void children(int l, size_t k, const uint64_t*& a, const uint64_t*& b) const {
    const size_t p = pos(l, k);
    a = elements + 4 * (p - ((size_t)1 << l));
    b = elements + 4 * (p - 1);
  }
  __device__ __forceinline__ Mmr for_set(unsigned) const { return *this; }
};

// What an append to an MMR of n0 leaves touches, compact: the n_peaks = popcount(n0) old peaks (largest mountain first),
// then the new elements (post-order positions s0 = mmr_size(n0) and up).  A new node's right child is always new; its
// left child is new or the old peak of that height (bit l - 1 of n0), whose slot is the number of set bits above it.
// Device memory for an append of m leaves is O(m + log n0), like the reference's add_leaf (:89-120), not O(n0).
struct MmrAppend {
  uint64_t* buf;
  size_t n0, s0;
  uint32_t n_peaks;
  __device__ __forceinline__ uint64_t* slot(size_t p) const { return buf + 4 * ((size_t)n_peaks + p - s0); }
  __device__ __forceinline__ uint64_t* at(int l, size_t k) const { return slot(Mmr::pos(l, k)); }
  __device__ __forceinline__ void children(int l, size_t k, const uint64_t*& a, const uint64_t*& b) const {
    const size_t p = Mmr::pos(l, k), pa = p - ((size_t)1 << l);
    a = pa >= s0 ? slot(pa) : buf + 4 * (size_t)__popcll((unsigned long long)(n0 >> l));
    b = slot(p - 1);
  }
  __device__ __forceinline__ MmrAppend for_set(unsigned) const { return *this; }
};

// The levels above a gathered array of subtree roots (the finish of a subtree-sharded tree): level 0 = the
// n_roots roots (read only, root_step digests apart: 1 for a dense array, the row length when they are one column of an
// all-gathered matrix), levels 1 .. log2(n_roots) level-major in `out` (n_roots/2, n_roots/4, ... digests).
// A batch of independent sets (blockIdx.y; the rounds of a sharded MMR) lies roots_stride / out_stride digests apart.
struct TopRoots {
  const uint64_t* roots;
  uint64_t* out;
  size_t n_roots, roots_stride, out_stride, root_step;
  __device__ __forceinline__ uint64_t* at(int l, size_t k) const {
    if (l == 0) return const_cast<uint64_t*>(roots) + 4 * k * root_step;
    return out + 4 * (n_roots - (n_roots >> (l - 1)) + k);
  }
  __device__ __forceinline__ void children(int l, size_t k, const uint64_t*& a, const uint64_t*& b) const {
    a = at(l - 1, 2 * k); b = at(l - 1, 2 * k + 1);
  }
  __device__ __forceinline__ TopRoots for_set(unsigned set) const {
    return TopRoots{roots + 4 * roots_stride * set, out + 4 * out_stride * set, n_roots, roots_stride, out_stride, root_step};
  }
  TopRoots for_set_host(unsigned set) const {
    return TopRoots{roots + 4 * roots_stride * set, out + 4 * out_stride * set, n_roots, 0, 0, root_step};
  }
};

// a plain array of digests (outputs of the Hasher batch calls): node (0, k) at out + 4 k
struct Flat {
  uint64_t* out;
  __device__ __forceinline__ uint64_t* at(int, size_t k) const { return out + 4 * k; }
  __device__ __forceinline__ Flat for_set(unsigned) const { return *this; }
};

// ---------------------------------------------------------------------------------------------------------------
// kernels.  Block = 128 threads: the permutation needs ~90 registers, so 5 blocks (20 warps) are resident per SM and
// grids are sized in whole waves of 148 x 5 blocks by the host where the level is large enough.
// ---------------------------------------------------------------------------------------------------------------
#ifndef PMT_BLOCK
#define PMT_BLOCK 128
#endif
#ifndef PMT_MINB_L1
#define PMT_MINB_L1 5
#endif
#ifndef PMT_MINB
#define PMT_MINB 5   // minimum resident blocks per SM asked of ptxas (register cap: 96 registers, no spills, 20 warps/SM;
                     // uncapped ptxas takes 132 = 12 warps/SM and 4 % less throughput); tuned with tools/ab_level.cu
#endif
constexpr int BLOCK = PMT_BLOCK;

// level 0: digest(0, k0 + i) = hash_or_noop(row i),  rows row-major count x w
template <class Layout>
__global__ void __launch_bounds__(BLOCK, PMT_MINB) k_leaves(Layout lay, const uint64_t* __restrict__ rows, size_t w, size_t k0,
                                                  size_t count) {
  for (size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x; i < count; i += (size_t)gridDim.x * BLOCK)
    store_digest(lay.at(0, k0 + i), hash_or_noop(rows + i * w, w));
}

// levels 0 AND 1 in one pass for narrow leaves (w <= 4: hash_or_noop is a canonicalising copy): thread k pads rows 2k and
// 2k + 1 into their digests, stores both and their parent.  Saves the separate copy pass (one read + one write of all
// leaf digests: 0.3 ms of the 13.7 ms of a 2^24-leaf tree).  count = number of level-1 nodes.
template <class Layout>
__global__ void __launch_bounds__(BLOCK, PMT_MINB_L1) k_leaves_level1(Layout lay, const uint64_t* __restrict__ rows, size_t w, size_t k0,
                                                                   size_t count) {
  for (size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x; i < count; i += (size_t)gridDim.x * BLOCK) {
    const size_t k = k0 + i;
    const Digest a = hash_or_noop(rows + (2 * i) * w, w), b = hash_or_noop(rows + (2 * i + 1) * w, w);
    store_digest(lay.at(0, 2 * k), a);
    store_digest(lay.at(0, 2 * k + 1), b);
    store_digest(lay.at(1, k), two_to_one(a, b));
  }
}

// level 0 straight from the prover's column-major LDE output ([UPSTREAM plonky2 fri module, PolynomialBatch::from_values /
// from_coeffs]: leaves = reverse_index_bits(transpose(columns)), i.e. leaf i = (col_0[rev(i)], ..., col_{w-1}[rev(i)])).
// Thread t reads element t of every column (coalesced) and owns leaf i = rev(t): the transpose and the bit reversal
// cost one scattered 32-byte digest store per leaf instead of a 1 GiB round trip through a row-major copy.
// rows_out (optional): the row-major leaves upstream's MerkleTree keeps for openings.
template <class Layout>
__global__ void __launch_bounds__(BLOCK, PMT_MINB) k_leaves_columns(Layout lay, const uint64_t* __restrict__ cols, size_t n, size_t w,
                                                          int log2n, bool bit_reverse, uint64_t* __restrict__ rows_out) {
  for (size_t t = (size_t)blockIdx.x * BLOCK + threadIdx.x; t < n; t += (size_t)gridDim.x * BLOCK) {
    const size_t i = (bit_reverse && log2n > 0) ? (size_t)(__brevll((unsigned long long)t) >> (64 - log2n)) : t;
    Digest d;
    if (w <= 4) {   // hash_or_noop: canonicalising copy
#pragma unroll
      for (int c = 0; c < 4; c++) d.v[c] = (size_t)c < w ? gl::canonical(cols[(size_t)c * n + t]) : 0ull;
      if (rows_out)
        for (size_t c = 0; c < w; c++) rows_out[i * w + c] = cols[c * n + t];
    } else {
      uint64_t s[WIDTH];
#pragma unroll
      for (int c = 0; c < WIDTH; c++) s[c] = 0;
      for (size_t off = 0; off < w; off += 8) {
#pragma unroll
        for (int c = 0; c < 8; c++)
          if (off + c < w) {
            s[c] = cols[(off + c) * n + t];
            if (rows_out) rows_out[i * w + off + c] = s[c];
          }
        permute(s);
      }
#pragma unroll
      for (int c = 0; c < 4; c++) d.v[c] = gl::canonical(s[c]);
    }
    store_digest(lay.at(0, i), d);
  }
}

// one level: digest(l, k) = two_to_one(children) for k in [k0, k0 + count), one node per thread
template <class Layout>
__global__ void __launch_bounds__(BLOCK, PMT_MINB) k_level(Layout lay, int l, size_t k0, size_t count) {
  for (size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x; i < count; i += (size_t)gridDim.x * BLOCK) {
    const uint64_t *a, *b;
    lay.children(l, k0 + i, a, b);
    store_digest(lay.at(l, k0 + i), two_to_one(load_digest(a), load_digest(b)));
  }
}

// SEVERAL consecutive big levels in ONE launch, as a wavefront.  One launch per level pays a drain at the end of every level
// (the last blocks run alone at a fraction of the machine's throughput: 11 - 22 us per level, profiles/launches_r2_summary.txt)
// although almost all of the next level's nodes have had their children for a long time.  Here the blocks of level la come
// first, then those of la + 1, ...: a block of level l + 1 owns 128 nodes whose children are exactly the level-l blocks 2 b and
// 2 b + 1 (level l starts at an even node: bit l of n0 is 0, checked by the host), waits for their two flags -- set long ago
// except at the very end of a level -- and runs while the level below drains.  The logical block number is taken from a
// counter when the block STARTS, so a block only ever waits for blocks that started before it: no deadlock whatever order the
// hardware dispatches in.  Flags carry the launch's epoch (never reset); the wait is bounded like k_exchange_top's.
// Level l has the nodes [n0 >> l, n1 >> l), as in launch_level_span.
// rows != nullptr: la == 1 and the first level is computed from the narrow leaf rows (w <= 4) themselves, as k_leaves_level1
// does: thread i pads rows 2 i and 2 i + 1 (relative to `rows`, the first leaf of the range) into their level-0 digests, stores
// both and hashes their parent -- the whole build of a big tree of narrow leaves below the cooperative tail is then ONE launch.
template <class Layout>
__global__ void __launch_bounds__(BLOCK, PMT_MINB) k_levels_wave(Layout lay, int la, int n_levels, size_t n0, size_t n1,
                                                                 unsigned* __restrict__ flags, unsigned epoch, unsigned* __restrict__ counter,
                                                                 long long timeout_cycles, unsigned* fault,
                                                                 const uint64_t* __restrict__ rows, size_t w) {
  __shared__ unsigned s_id;
  if (threadIdx.x == 0) {
    s_id = atomicAdd(counter, 1u);
    if (s_id == gridDim.x - 1) *counter = 0;             // every number has been handed out: ready for the next launch
  }
  __syncthreads();
  size_t id = s_id, off = 0, prev_blocks = 0, k0 = 0, cnt = 0;
  int j = 0;
  for (;; j++) {                                         // which level this block belongs to, and which block of it
    k0 = n0 >> (la + j);
    cnt = (n1 >> (la + j)) - k0;
    const size_t nb = (cnt + BLOCK - 1) / BLOCK;
    if (id < nb || j + 1 == n_levels) break;
    id -= nb;
    off += nb;
    prev_blocks = nb;
  }
  if (j > 0) {                                           // the two blocks of the level below that hold this block's children
    if (threadIdx.x < 2) {
      const size_t child = 2 * id + threadIdx.x;
      if (child < prev_blocks) {
        const volatile unsigned* f = flags + (off - prev_blocks) + child;
        const long long t0 = clock64();
        while (*f != epoch) {
          if (clock64() - t0 > timeout_cycles) { *reinterpret_cast<volatile unsigned*>(fault) = 0x80000000u; break; }
          __nanosleep(32);
        }
      }
      __threadfence();
    }
    __syncthreads();
  }
  const size_t i = id * BLOCK + threadIdx.x;
  if (i < cnt) {
    Digest l, r;
    if (rows != nullptr && j == 0) {                     // leaf rows -> level 0 (a canonicalising copy) -> level 1
      l = hash_or_noop(rows + (2 * i) * w, w);
      r = hash_or_noop(rows + (2 * i + 1) * w, w);
      store_digest(lay.at(0, 2 * (k0 + i)), l);
      store_digest(lay.at(0, 2 * (k0 + i) + 1), r);
    } else {
      const uint64_t *a, *b;
      lay.children(la + j, k0 + i, a, b);                // ld.cg: the children may have been written by another SM in this launch
#pragma unroll
      for (int q = 0; q < 4; q++) { l.v[q] = __ldcg(a + q); r.v[q] = __ldcg(b + q); }
    }
    store_digest(lay.at(la + j, k0 + i), two_to_one(l, r));
  }
  if (j + 1 < n_levels) {                                // publish: every thread's digests device-wide, then the block's flag
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned*>(flags + off + id) = epoch;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// cooperative kernels (several threads per permutation, poseidon_coop.cuh) for everything that is a chain of dependent
// permutations: levels too small to fill the GPU with one thread per node, proof paths, sponges over few rows.
// Two forms (Form = Quad: 4 threads per state, throughput; Wide: 16 lanes per state, latency), one kernel body.
// Block = 256 threads = 8 warps = 64 Quad states or 16 Wide states in flight.
// ---------------------------------------------------------------------------------------------------------------
constexpr int COOP_BLOCK = 256;
constexpr int COOP_WARPS = COOP_BLOCK / 32;
constexpr int COOP_NODES = COOP_BLOCK / 4;     // Quad states a block holds at once
constexpr int WIDE_NODES = COOP_BLOCK / 16;    // Wide states a block holds at once
constexpr int COOP_LOCAL_LEVELS = 7;           // 64, 32, 16, 8, 4, 2, 1 nodes: the subtree a block reduces on its own
constexpr unsigned TICKET_GROUP = 8;           // blocks per first-round ticket group of k_tree_coop
constexpr int TICKET_GROUP_LEVELS = 3;         // log2(TICKET_GROUP): the levels a group's finisher computes
using CoopShared = poseidon::coop::Shared<COOP_WARPS>;
using poseidon::coop::Quad;
using poseidon::coop::Wide;

// one two_to_one by one group of threads; groups without a node (`active` false) still run the permutation's warp-wide
// operations.  Children come from shared memory when this block produced them at the previous level (`ch`: the two
// digests, 8 consecutive words), else from global memory with ld.global.cg (another SM may have written them earlier in the
// same launch).  The digest goes to its place in the layout -- every level is stored once, proofs need it -- and, if `keep`
// is given, to shared memory for the next level: the parent does not wait for an L2 round trip.
template <class Form, class Layout>
__device__ __forceinline__ void coop_node(const Layout& lay, int l, size_t k, bool active, const uint64_t* ch, uint64_t* keep,
                                          const Form& t, const CoopShared& sh) {
  uint64_t e[Form::ELEMS];
  const uint64_t *a = nullptr, *b = nullptr;
  if (active && !ch) lay.children(l, k, a, b);
#pragma unroll
  for (int i = 0; i < Form::ELEMS; i++) {
    const unsigned idx = t.elem(i);
    e[i] = 0ull;
    if (active && idx < 8) e[i] = ch ? ch[idx] : __ldcg(idx < 4 ? a + idx : b + (idx - 4));
  }
  t.permute(e, sh);
#pragma unroll
  for (int i = 0; i < Form::ELEMS; i++) {
    const unsigned idx = t.elem(i);
    if (active && idx < 4) {
      const uint64_t d = gl::canonical(e[i]);
      lay.at(l, k)[idx] = d;
      if (keep) keep[idx] = d;
    }
  }
}

// Digests a block hands from one level to the next without leaving the SM: the nodes of a level alternate between two
// regions (64 and 32 digests), node i of the level below at words 4 i of the other region.
constexpr int KEEP_A = 64, KEEP_B = 32;
struct KeepBuf {
  uint64_t w[(KEEP_A + KEEP_B) * 4];
  __device__ __forceinline__ uint64_t* region(int parity) { return parity ? w + 4 * KEEP_A : w; }
  __device__ __forceinline__ static size_t room(int parity) { return parity ? KEEP_B : KEEP_A; }
};

// the nodes k_first + [0, nb) of level l by one block: in passes of 64 by quads, or -- when there are at most 16, where
// only latency counts -- in one pass by 16-lane groups.  Warps all of whose groups have no node skip the permutation
// (its exchanges never cross a warp).  nb is the same for all threads of the block.  from: the level below as this block
// kept it (node i's children at digests 2 i, 2 i + 1), or nullptr; to: where to keep this level (room for nb), or nullptr.
template <class Layout>
__device__ __forceinline__ void coop_level(const Layout& lay, int l, size_t k_first, size_t nb, const uint64_t* from, uint64_t* to,
                                           const Quad& q, const Wide& w, const CoopShared& sh) {
  const unsigned warp = threadIdx.x >> 5;
  // measured per level (tools/perm_bench.cu lat, profiles/latency_r2.jsonl): Quad 12.9 / 9.3 / 8.8 us for 64 / 32 / <= 16 nodes
  // of a block (8 / 4 / <= 2 warps), Wide 8.4 us for 16 nodes (8 warps) and 6.3 us for <= 8 (one warp per sub-partition);
  // switching at 8 instead of 16 nodes made the tail of a tree 9 us slower (profiles/tail_r2.jsonl)
  if (nb > (size_t)WIDE_NODES) {
    for (size_t base = 0; base < nb; base += COOP_NODES)
      if (base + warp * 8 < nb) {
        const size_t i = base + q.state();
        coop_node(lay, l, k_first + i, i < nb, from ? from + 8 * i : nullptr, to ? to + 4 * i : nullptr, q, sh);
      }
  } else if (warp * 2 < nb) {
    const size_t i = w.state();
    coop_node(lay, l, k_first + i, i < nb, from ? from + 8 * i : nullptr, to ? to + 4 * i : nullptr, w, sh);
  }
}

// The latency-bound part of a tree in ONE launch.  Level l0 has the nodes k0 + [0, count0); block b owns the 64 nodes
// k0 + [64 b, 64 b + 64) of it and every ancestor of theirs for `local_levels` levels (64, 32, ..., 1 nodes; needs k0 and
// count0 to be multiples of 2^(local_levels - 1)), a block barrier between levels: a block only reads children it wrote
// itself -- it keeps them in shared memory -- and every level is still stored once (proofs need it).  Then, for a perfect
// subtree (count0 a power of two, k0 a multiple of it), the LAST block to finish -- a ticket taken after a device-wide fence
// -- continues with the `top_levels` levels above the block roots (count0 >> local_levels nodes, halving), so the whole tail
// of the tree is one launch instead of three and there is no grid-wide barrier.  local_levels = 1, top_levels = 0: a plain level.
// Launches with top_levels > 0 have gridDim.y == 1 and need 1 + gridDim.x / 8 zeroed ticket counters.
template <class Layout>
__global__ void __launch_bounds__(COOP_BLOCK) k_tree_coop(Layout lay_in, int l0, size_t k0, size_t count0, int local_levels,
                                                           int top_levels, unsigned* __restrict__ ticket) {
  __shared__ CoopShared sh;
  __shared__ KeepBuf kept;
  __shared__ bool is_last;
  const Layout lay = lay_in.for_set(blockIdx.y);      // gridDim.y > 1 (batched finishes): one block per set, no tickets
  const Quad q = Quad::make(sh);
  const Wide w = Wide::make(sh);
  poseidon::coop::stage(sh);
  for (int j = 0; j < local_levels; j++) {
    const size_t per = (size_t)COOP_NODES >> j, cnt = count0 >> j;
    const size_t base = ((size_t)blockIdx.x * COOP_NODES) >> j;
    const size_t nb = base >= cnt ? 0 : (cnt - base < per ? cnt - base : per);
    coop_level(lay, l0 + j, (k0 >> j) + base, nb, j ? kept.region((j - 1) & 1) : nullptr,
               j + 1 < local_levels ? kept.region(j & 1) : nullptr, q, w, sh);
    if (j + 1 < local_levels) __syncthreads();
  }
  if (top_levels <= 0) return;
  // Tickets.  A finished block publishes its root (device-wide fence) and takes a ticket; whoever takes the LAST ticket of a
  // set of blocks continues with the levels above their roots.  With more than 8 blocks there are two rounds: the blocks of
  // a group of 8 elect a finisher for the 3 levels above their 8 roots (4, 2, 1 nodes: three Wide permutations on that
  // block's SM, 16 groups side by side), the groups then elect the one block that finishes the rest -- instead of one block
  // walking 64, 32, 16, ... nodes alone (12.9 + 9.3 + 8.4 us for the first three levels against 3 x 6.3).
  // ticket[0]: the final round; ticket[1 + group]: the groups.
  const unsigned blocks = gridDim.x;
  const bool two_rounds = blocks > TICKET_GROUP && top_levels > TICKET_GROUP_LEVELS;
  int done = 0;                                   // top levels finished so far on the path this block is on
  bool have = false;
  int parity = 0;
  if (two_rounds) {
    const unsigned grp = blockIdx.x / TICKET_GROUP;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned got = atomicAdd(ticket + 1 + grp, 1u);
      is_last = got == TICKET_GROUP - 1;          // blocks is a power of two > 8: every group is full
      if (is_last) ticket[1 + grp] = 0;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    size_t cnt = TICKET_GROUP / 2;
    for (; done < TICKET_GROUP_LEVELS; done++, cnt >>= 1) {
      const int j = local_levels + done;
      coop_level(lay, l0 + j, (k0 >> j) + (size_t)grp * cnt, cnt, have ? kept.region(parity ^ 1) : nullptr, kept.region(parity), q, w, sh);
      have = true;
      parity ^= 1;
      __syncthreads();
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned got = atomicAdd(ticket, 1u);
    is_last = got == (two_rounds ? blocks / TICKET_GROUP : blocks) - 1;
    if (is_last) ticket[0] = 0;                   // ready for the next launch that is handed this slot
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  size_t cnt = count0 >> (local_levels + done);
  have = false;                                   // the level below was written by other blocks: read it from global memory
  for (; done < top_levels; done++, cnt >>= 1) {
    const int j = local_levels + done;
    const bool fits = cnt <= KeepBuf::room(parity);
    coop_level(lay, l0 + j, k0 >> j, cnt, have ? kept.region(parity ^ 1) : nullptr, fits ? kept.region(parity) : nullptr, q, w, sh);
    have = fits;
    parity ^= 1;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The exchange of a subtree-sharded build FUSED with the levels above it, over peer memory ("direct P2P stores
// + flag").  Every rank owns a mailbox in its own HBM that its peers can write (one process: peer access; one process per
// GPU: CUDA IPC mappings made by pmt_comm_init): MAIL_RING slots x MAIL_MAX_WORLD ranks x MAIL_DIGESTS digests + one flag
// per (slot, rank).  Exchange number `seq` uses slot seq % MAIL_RING and the flag value seq + 1.
//   push   thread p of block (0, 0) stores this rank's digests into peer p's mailbox (NVLink stores), fences system-wide, then
//          stores the flag: a peer that sees the flag sees the digests
//   wait   thread p of every block spins on the flag of rank p in its OWN mailbox (local memory, volatile loads)
//   top    block y finishes set y: the G gathered roots of column y -> log2 G levels (Wide / Quad, digests handed on in shared
//          memory) -> `tops`, and the last level's nodes -> `finals`
// so the step that used to be ncclAllGather + a finish launch + a copy is ONE launch, and the transfer is 32 bytes per peer
// written straight from the kernel.  A rank cannot get more than one exchange ahead of the slowest peer (it waits for all
// peers' flags of exchange s before it leaves s), so a ring of 4 slots is never overwritten while in use.  The spin is bounded
// (timeout_cycles): on expiry the block raises *fault (host-mapped) and pmt_sync reports it instead of hanging the device.
// ---------------------------------------------------------------------------------------------------------------
constexpr unsigned MAIL_RING = 4, MAIL_MAX_WORLD = 64, MAIL_DIGESTS = 64;
constexpr size_t MAIL_DATA_WORDS = (size_t)MAIL_RING * MAIL_MAX_WORLD * MAIL_DIGESTS * 4;
constexpr size_t MAIL_BYTES = MAIL_DATA_WORDS * 8 + (size_t)MAIL_RING * MAIL_MAX_WORLD * 4;
struct Exchange {
  uint64_t* const* peers;        // device array [world]: base of every rank's mailbox as THIS device addresses it; peers[rank] = own
  unsigned world, rank, seq;
  long long timeout_cycles;
  unsigned* fault;               // host-mapped word, set when a wait times out
  __device__ __forceinline__ static uint64_t* data(uint64_t* base, unsigned slot, unsigned from) {
    return base + ((size_t)slot * MAIL_MAX_WORLD + from) * MAIL_DIGESTS * 4;
  }
  __device__ __forceinline__ static unsigned* flag(uint64_t* base, unsigned slot, unsigned from) {
    return reinterpret_cast<unsigned*>(base + MAIL_DATA_WORDS) + (size_t)slot * MAIL_MAX_WORLD + from;
  }
};

// mine: this rank's n_mine digests (<= MAIL_DIGESTS).  gathered: world x n_mine digests, row p = rank p's (an output the callers
// keep: d_roots / d_cap / the MMR's gathered matrix).  sets = gridDim.y sets of `world` roots = columns 0 .. sets - 1 of
// `gathered` (sets = 0: gather only, gridDim.y = 1); set y's levels go to tops + 4 (world - 1) y (level-major, as TopRoots), the
// n_final = world >> top_levels nodes of its last level to finals + 4 n_final y.  Block 0 also moves the columns that belong
// to no set: into `gathered`, and the last rank's into finals + 4 n_final sets (the peaks of a sharded MMR's tail).
__global__ void __launch_bounds__(COOP_BLOCK) k_exchange_top(Exchange x, const uint64_t* mine, unsigned n_mine, uint64_t* gathered,
                                                              unsigned sets, uint64_t* tops, int top_levels, uint64_t* finals) {
  // no __restrict__: `mine` is the rank's own row of `gathered` in the one-process-per-GPU forms
  __shared__ CoopShared sh;
  __shared__ KeepBuf kept;
  const unsigned t = threadIdx.x, y = blockIdx.y, G = x.world, slot = x.seq % MAIL_RING, epoch = x.seq + 1u;
  uint64_t* const own = x.peers[x.rank];
  if (y == 0 && t < G && t != x.rank) {                       // push: one thread per peer
    uint64_t* dst = Exchange::data(x.peers[t], slot, x.rank);
    for (unsigned i = 0; i < 4 * n_mine; i++) dst[i] = mine[i];
    __threadfence_system();
    *reinterpret_cast<volatile unsigned*>(Exchange::flag(x.peers[t], slot, x.rank)) = epoch;
  }
  if (top_levels > 0) poseidon::coop::stage(sh);              // the tables load while the peers' stores are in flight
  if (t < G && t != x.rank) {                                 // wait: one thread per peer
    const volatile unsigned* f = Exchange::flag(own, slot, t);
    const long long t0 = clock64();
    while (*f != epoch) {
      if (clock64() - t0 > x.timeout_cycles) { *reinterpret_cast<volatile unsigned*>(x.fault) = 1u + t; break; }
      __nanosleep(64);
    }
    __threadfence_system();
  }
  __syncthreads();
  // this block's columns of the gathered matrix: column y (its set), and for block 0 every column outside the sets
  const unsigned n_final = top_levels >= 0 ? G >> top_levels : 0;
  for (unsigned i = t; i < G * 4 * n_mine; i += COOP_BLOCK) {
    const unsigned p = i / (4 * n_mine), col = (i / 4) % n_mine, wd = i % 4;
    const bool in_set = col < sets, my_col = in_set ? col == y : y == 0;
    if (!my_col) continue;
    const uint64_t v = p == x.rank ? mine[4 * col + wd] : __ldcv(Exchange::data(own, slot, p) + 4 * col + wd);
    gathered[((size_t)p * n_mine + col) * 4 + wd] = v;
    if (in_set) kept.region(0)[4 * p + wd] = v;
    else if (finals && p == G - 1) finals[4 * ((size_t)n_final * sets + (col - sets)) + wd] = v;
  }
  if (sets == 0 || top_levels <= 0) return;
  __syncthreads();
  const Quad q = Quad::make(sh);
  const Wide w = Wide::make(sh);
  const TopRoots lay{gathered + 4 * y, tops + 4 * (size_t)(G - 1) * y, G, 0, 0, n_mine};
  int parity = 1;
  for (int j = 1; j <= top_levels; j++, parity ^= 1) {
    coop_level(lay, j, 0, (size_t)(G >> j), kept.region(parity ^ 1), kept.region(parity), q, w, sh);
    __syncthreads();
  }
  if (finals)
    for (unsigned i = t; i < 4 * n_final; i += COOP_BLOCK) finals[4 * (size_t)n_final * y + i] = kept.region(parity ^ 1)[i];
}

// level 0 by groups of threads: digest(0, k0 + i) = hash_or_noop(row i) (NOOP_RULE) or hash_no_pad(row i).  The sponge's
// permutations are sequential, so for few rows (a small FRI commitment, one bag of peaks) the cooperative forms cut the
// latency 3 - 6x; the owner of element idx < 8 reads felt off + idx of every 8-felt block.  blockDim.x / Form::LANES rows
// per block and pass.
template <class Layout, bool NOOP_RULE, class Form>
__global__ void __launch_bounds__(COOP_BLOCK) k_rows_coop(Layout lay, const uint64_t* __restrict__ rows, size_t w, size_t k0, size_t count) {
  __shared__ CoopShared sh;
  const Form t = Form::make(sh);
  poseidon::coop::stage(sh);
  const size_t per_block = blockDim.x / Form::LANES, warp_first = (threadIdx.x >> 5) * (32 / Form::LANES);
  for (size_t base = (size_t)blockIdx.x * per_block; base < count; base += (size_t)gridDim.x * per_block) {
    if (base + warp_first >= count) continue;                  // warp-uniform
    const size_t i = base + t.state();
    const bool active = i < count;
    const uint64_t* row = rows + (active ? i : 0) * w;
    uint64_t e[Form::ELEMS];
#pragma unroll
    for (int a = 0; a < Form::ELEMS; a++) e[a] = 0;
    if (NOOP_RULE && w <= 4) {
#pragma unroll
      for (int a = 0; a < Form::ELEMS; a++)
        if (active && t.elem(a) < w) e[a] = row[t.elem(a)];
    } else {
      for (size_t off = 0; off < w; off += 8) {                // overwrite mode: lanes past the row's end keep their state
#pragma unroll
        for (int a = 0; a < Form::ELEMS; a++) {
          const unsigned idx = t.elem(a);
          if (active && idx < 8 && off + idx < w) e[a] = row[off + idx];
        }
        t.permute(e, sh);
      }
    }
#pragma unroll
    for (int a = 0; a < Form::ELEMS; a++)
      if (active && t.elem(a) < 4) lay.at(0, k0 + i)[t.elem(a)] = gl::canonical(e[a]);
  }
}

// generic batches (parity hooks of the Hasher trait)
__global__ void __launch_bounds__(BLOCK, PMT_MINB) k_permute(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += (size_t)gridDim.x * BLOCK) {
    uint64_t s[WIDTH];
#pragma unroll
    for (int j = 0; j < WIDTH; j++) s[j] = in[i * WIDTH + j];
    permute(s);
#pragma unroll
    for (int j = 0; j < WIDTH; j++) out[i * WIDTH + j] = gl::canonical(s[j]);
  }
}

// the same by quads (small batches): all 12 output elements
__global__ void __launch_bounds__(COOP_BLOCK) k_permute_coop(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, size_t n) {
  __shared__ CoopShared sh;
  const Quad t = Quad::make(sh);
  poseidon::coop::stage(sh);
  const size_t warp_first = (threadIdx.x >> 5) * 8;
  for (size_t base = (size_t)blockIdx.x * COOP_NODES; base < n; base += (size_t)gridDim.x * COOP_NODES) {
    if (base + warp_first >= n) continue;
    const size_t i = base + t.state();
    uint64_t e[3] = {0, 0, 0};
    if (i < n)
#pragma unroll
      for (int a = 0; a < 3; a++) e[a] = in[i * WIDTH + t.elem(a)];
    t.permute(e, sh);
    if (i < n)
#pragma unroll
      for (int a = 0; a < 3; a++) out[i * WIDTH + t.elem(a)] = gl::canonical(e[a]);
  }
}

// out[i] = two_to_one(l[i], r[i]); `stride` u64 between consecutive inputs (4 = dense arrays, 8 = adjacent sibling pairs)
__global__ void __launch_bounds__(BLOCK, PMT_MINB) k_two_to_one(const uint64_t* __restrict__ l, const uint64_t* __restrict__ r,
                                                      uint64_t* __restrict__ out, size_t n, size_t stride) {
  for (size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += (size_t)gridDim.x * BLOCK)
    store_digest(out + 4 * i, two_to_one(load_digest(l + stride * i), load_digest(r + stride * i)));
}

template <bool NOOP_RULE>
__global__ void __launch_bounds__(BLOCK, PMT_MINB) k_hash_rows(const uint64_t* __restrict__ rows, size_t n, size_t w,
                                                     uint64_t* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += (size_t)gridDim.x * BLOCK)
    store_digest(out + 4 * i, NOOP_RULE ? hash_or_noop(rows + i * w, w) : hash_no_pad(rows + i * w, w));
}

// ---------------------------------------------------------------------------------------------------------------
// proofs: pure gathers, one thread per (proof, level).  An index the reference would panic on (assert!
// simple_merkle_tree.rs:56, :77) yields an all-zero proof here; the host-buffer forms and the mirrors reject it first.
// ---------------------------------------------------------------------------------------------------------------
// simple_merkle_tree.rs:55-74 get_merkle_proof: level_i[idx_i ^ 1], i = 0 .. log2(n) - 1
__global__ void k_simple_prove(LevelMajor lay, const uint64_t* __restrict__ idx, size_t n_idx, uint64_t* __restrict__ out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int depth = lay.top;
  if (t >= n_idx * (size_t)depth) return;
  const size_t q = t / depth; const int l = (int)(t % depth);
  Digest d = {{0, 0, 0, 0}};
  if (idx[q] < lay.n) d = load_digest(lay.at(l, (idx[q] >> l) ^ 1));
  store_digest(out + 4 * t, d);
}

// [UPSTREAM hash/merkle_tree.rs MerkleTree::prove]
__global__ void k_plonky2_prove(Plonky2 lay, size_t n, const uint64_t* __restrict__ idx, size_t n_idx, uint64_t* __restrict__ out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int depth = lay.sub_levels;
  if (t >= n_idx * (size_t)depth) return;
  const size_t q = t / depth; const int l = (int)(t % depth);
  Digest d = {{0, 0, 0, 0}};
  if (idx[q] < n) d = load_digest(lay.at(l, (idx[q] >> l) ^ 1));
  store_digest(out + 4 * t, d);
}

// merkle_mountain_ranges.rs:147-176 get_subtree_proof_elm, closed form: the leaf's mountain has height H = the bit of
// n_leaves that covers it; entry j = (node (j, (i >> j) ^ 1), sibling_on_left = bit j of i).  i >= n_leaves: path_len 0.
__global__ void k_mmr_prove(Mmr lay, size_t n_leaves, const uint64_t* __restrict__ idx, size_t n_idx,
                            uint64_t* __restrict__ sib_out, uint8_t* __restrict__ left_out, uint32_t* __restrict__ len_out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_idx * 32) return;
  const size_t q = t / 32; const int j = (int)(t % 32);
  const size_t i = idx[q];
  // mountains from the largest: the leaf is in the first mountain whose end exceeds i
  int H = 0; size_t base = 0;
  if (i < n_leaves)
    for (int b = 63 - __clzll((unsigned long long)n_leaves); b >= 0; b--) {
      if ((n_leaves >> b) & 1) {
        if (i < base + ((size_t)1 << b)) { H = b; break; }
        base += (size_t)1 << b;
      }
    }
  if (j == 0) len_out[q] = (uint32_t)H;
  if (j < H) {
    store_digest(sib_out + 4 * t, load_digest(lay.at(j, (i >> j) ^ 1)));
    left_out[t] = (uint8_t)((i >> j) & 1);
  }
}

// merkle_mountain_ranges.rs:179-200 get_peaks: one peak per set bit of n_leaves, largest mountain first
__global__ void k_mmr_peaks(Mmr lay, size_t n_leaves, uint64_t* __restrict__ out) {
  const int want = threadIdx.x;
  int seen = 0; size_t base = 0;
  for (int b = 63 - __clzll((unsigned long long)n_leaves); b >= 0; b--) {
    if ((n_leaves >> b) & 1) {
      if (seen == want) { store_digest(out + 4 * want, load_digest(lay.at(b, base >> b))); return; }
      seen++; base += (size_t)1 << b;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// verification, big batches: one thread folds one path
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool digest_eq(const Digest& a, const Digest& b) {
  return a.v[0] == b.v[0] && a.v[1] == b.v[1] && a.v[2] == b.v[2] && a.v[3] == b.v[3];
}
__device__ __forceinline__ Digest load_digest_canonical(const uint64_t* __restrict__ p) {
  Digest d = load_digest(p);
#pragma unroll
  for (int i = 0; i < 4; i++) d.v[i] = gl::canonical(d.v[i]);
  return d;
}
constexpr uint32_t MMR_MAX_PATH = 32;   // entries per proof in the batch layout (an MMR has < 2^30 leaves, :264)

// simple_merkle_tree.rs:91-109 verify_merkle_proof and [UPSTREAM hash/merkle_proofs.rs verify_merkle_proof_to_cap]:
// fold by index parity, compare with cap[index >> path_len].  idx_mask: the simple tree's verifier only ever looks at
// the parities of the low path_len index bits (:97-105) and has one root, so its index is masked to those bits; upstream's
// compares with cap[index >> path_len] and gets all ones.
__global__ void __launch_bounds__(BLOCK, PMT_MINB) k_verify_to_cap(const uint64_t* __restrict__ rows, size_t w,
                                                         const uint64_t* __restrict__ idx, size_t idx_mask, size_t n_idx,
                                                         const uint64_t* __restrict__ cap, uint32_t cap_height,
                                                         const uint64_t* __restrict__ proofs, size_t path_len,
                                                         uint8_t* __restrict__ ok) {
  const size_t q = (size_t)blockIdx.x * BLOCK + threadIdx.x;
  if (q >= n_idx) return;
  size_t index = idx[q] & idx_mask;
  Digest cur = hash_or_noop(rows + q * w, w);
  for (size_t j = 0; j < path_len; j++) {
    const Digest sib = load_digest(proofs + 4 * (q * path_len + j));
    cur = (index & 1) ? two_to_one(sib, cur) : two_to_one(cur, sib);
    index >>= 1;
  }
  ok[q] = (index < ((size_t)1 << cap_height)) && digest_eq(cur, load_digest_canonical(cap + 4 * index));
}

// merkle_mountain_ranges.rs:232-252 MMR_proof::verify: fold by sibling_on_left, membership in peaks (else the
// reference panics: status -1), re-bag, compare with root.  A path longer than the batch layout holds is -1 as well.
__global__ void __launch_bounds__(BLOCK, PMT_MINB) k_mmr_verify(const uint64_t* __restrict__ leaves, size_t n_idx,
                                                      const uint64_t* __restrict__ sib, const uint8_t* __restrict__ left,
                                                      const uint32_t* __restrict__ len, const uint64_t* __restrict__ peaks,
                                                      uint32_t n_peaks, const uint64_t* __restrict__ bagged,
                                                      const uint64_t* __restrict__ root, int8_t* __restrict__ status) {
  const size_t q = (size_t)blockIdx.x * BLOCK + threadIdx.x;
  if (q >= n_idx) return;
  const uint32_t L = len[q];
  if (L > MMR_MAX_PATH) { status[q] = -1; return; }
  Digest cur = hash_or_noop(leaves + q, 1);
  for (uint32_t j = 0; j < L; j++) {
    const Digest s = load_digest(sib + 4 * (q * MMR_MAX_PATH + j));
    cur = left[q * MMR_MAX_PATH + j] ? two_to_one(s, cur) : two_to_one(cur, s);
  }
  bool found = false;
  for (uint32_t k = 0; k < n_peaks; k++) found |= digest_eq(cur, load_digest_canonical(peaks + 4 * k));
  if (!found) { status[q] = -1; return; }
  // the re-bagged root is identical for every proof of the batch: computed once (k_rows_coop) into `bagged`
  status[q] = digest_eq(load_digest(bagged), load_digest_canonical(root)) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------------
// cooperative verification (batches too small to fill the GPU with one thread per proof): a group of threads folds one
// path (Form = Wide for batches where only the latency of the chain counts, Quad for bigger ones).  The owners of
// elements 0..3 carry the running digest; placing it on the right-hand side of two_to_one moves it to elements 4..7 -- a
// register move in the quad form (element 4 + j lives in thread j), one shuffle pair in the 16-lane form.
// blockDim.x / Form::LANES proofs per block.
// ---------------------------------------------------------------------------------------------------------------
// element idx of the running digest (idx < 4; others 0) as held by the owner of element idx, and the same digest as seen
// by the owners of elements 4..7 (idx - 4)
__device__ __forceinline__ uint64_t digest_shifted(const Quad&, const uint64_t (&cur)[3]) { return cur[0]; }   // e[1] <- e[0]
__device__ __forceinline__ uint64_t digest_shifted(const Wide& t, const uint64_t (&cur)[1]) {                    // lane g <- lane g - 4
  const unsigned src = t.lane0 + ((t.g - 4) & 15);
  return gl::pack(__shfl_sync(0xffffffffu, gl::lo32(cur[0]), src), __shfl_sync(0xffffffffu, gl::hi32(cur[0]), src));
}
// one fold step: digest <- two_to_one(sibling, digest) if on_left else two_to_one(digest, sibling).  `sib` = the sibling
// digest (or nullptr: zeros).  cur holds the digest in the owners of elements 0..3 before and after.
template <class Form>
__device__ __forceinline__ void coop_fold(uint64_t (&cur)[Form::ELEMS], const uint64_t* __restrict__ sib, bool on_left, const Form& t,
                                          const CoopShared& sh) {
  const uint64_t moved = digest_shifted(t, cur);
  uint64_t e[Form::ELEMS];
#pragma unroll
  for (int a = 0; a < Form::ELEMS; a++) {
    const unsigned idx = t.elem(a);
    const uint64_t s_lo = (sib && idx < 4) ? sib[idx] : 0ull, s_hi = (sib && idx >= 4 && idx < 8) ? sib[idx - 4] : 0ull;
    e[a] = idx < 4 ? (on_left ? s_lo : cur[a]) : idx < 8 ? (on_left ? moved : s_hi) : 0ull;
  }
  t.permute(e, sh);
#pragma unroll
  for (int a = 0; a < Form::ELEMS; a++) cur[a] = e[a];
}

// merkle_mountain_ranges.rs:232-252, same contract as k_mmr_verify
template <class Form>
__global__ void __launch_bounds__(COOP_BLOCK) k_mmr_verify_coop(const uint64_t* __restrict__ leaves, size_t n_idx,
                                                                const uint64_t* __restrict__ sib, const uint8_t* __restrict__ left,
                                                                const uint32_t* __restrict__ len, const uint64_t* __restrict__ peaks,
                                                                uint32_t n_peaks, const uint64_t* __restrict__ bagged,
                                                                const uint64_t* __restrict__ root, int8_t* __restrict__ status) {
  __shared__ CoopShared sh;
  const Form t = Form::make(sh);
  poseidon::coop::stage(sh);
  const size_t q = (size_t)blockIdx.x * (blockDim.x / Form::LANES) + t.state();
  const bool have = q < n_idx;
  const uint32_t L_raw = have ? len[q] : 0;
  const bool too_long = L_raw > MMR_MAX_PATH;
  const uint32_t L = too_long ? 0 : L_raw;
  const uint32_t Lw = __reduce_max_sync(0xffffffffu, L);              // every group of the warp runs the same trip count
  uint64_t cur[Form::ELEMS];
#pragma unroll
  for (int a = 0; a < Form::ELEMS; a++) cur[a] = (have && t.elem(a) == 0) ? leaves[q] : 0ull;   // hash_or_noop(&[leaf]) = [leaf, 0, 0, 0]
  for (uint32_t j = 0; j < Lw; j++) {
    const bool act = j < L;
    uint64_t nxt[Form::ELEMS];
#pragma unroll
    for (int a = 0; a < Form::ELEMS; a++) nxt[a] = cur[a];
    coop_fold(nxt, act ? sib + 4 * (q * MMR_MAX_PATH + j) : nullptr, act && left[q * MMR_MAX_PATH + j] != 0, t, sh);
#pragma unroll
    for (int a = 0; a < Form::ELEMS; a++) cur[a] = act ? nxt[a] : cur[a];
  }
  bool found = false;
  for (uint32_t k = 0; k < n_peaks; k++) {
    bool eq = true;
#pragma unroll
    for (int a = 0; a < Form::ELEMS; a++)
      if (t.elem(a) < 4) eq = eq && gl::canonical(cur[a]) == gl::canonical(peaks[4 * k + t.elem(a)]);
    found |= t.all(eq);
  }
  if (have && t.elem(0) == 0) {
    if (!found || too_long) status[q] = -1;
    else status[q] = digest_eq(load_digest(bagged), load_digest_canonical(root)) ? 1 : 0;
  }
}

// simple_merkle_tree.rs:91-109 / [UPSTREAM hash/merkle_proofs.rs verify_merkle_proof_to_cap], same contract as k_verify_to_cap
template <class Form>
__global__ void __launch_bounds__(COOP_BLOCK) k_verify_to_cap_coop(const uint64_t* __restrict__ rows, size_t w,
                                                                   const uint64_t* __restrict__ idx, size_t idx_mask, size_t n_idx,
                                                                   const uint64_t* __restrict__ cap, uint32_t cap_height,
                                                                   const uint64_t* __restrict__ proofs, size_t path_len,
                                                                   uint8_t* __restrict__ ok) {
  __shared__ CoopShared sh;
  const Form t = Form::make(sh);
  poseidon::coop::stage(sh);
  const size_t q = (size_t)blockIdx.x * (blockDim.x / Form::LANES) + t.state();
  const bool have = q < n_idx;
  const uint64_t* row = rows + (have ? q : 0) * w;
  uint64_t cur[Form::ELEMS];
#pragma unroll
  for (int a = 0; a < Form::ELEMS; a++) cur[a] = 0;
  if (w <= 4) {
#pragma unroll
    for (int a = 0; a < Form::ELEMS; a++)
      if (have && t.elem(a) < w) cur[a] = row[t.elem(a)];
  } else {
    for (size_t off = 0; off < w; off += 8) {            // overwrite-mode sponge; w is uniform, so is the trip count
#pragma unroll
      for (int a = 0; a < Form::ELEMS; a++) {
        const unsigned e_idx = t.elem(a);
        if (have && e_idx < 8 && off + e_idx < w) cur[a] = row[off + e_idx];
      }
      t.permute(cur, sh);
    }
#pragma unroll
    for (int a = 0; a < Form::ELEMS; a++)
      if (t.elem(a) >= 4) cur[a] = 0;                    // the digest is elements 0..3
  }
  size_t index = have ? idx[q] & idx_mask : 0;
  const uint64_t* pr = proofs + 4 * (have ? q : 0) * path_len;
  for (size_t j = 0; j < path_len; j++) {
    coop_fold(cur, have ? pr + 4 * j : nullptr, (index & 1) != 0, t, sh);
    index >>= 1;
  }
  const bool in_cap = index < ((size_t)1 << cap_height);
  bool eq = in_cap && have;
#pragma unroll
  for (int a = 0; a < Form::ELEMS; a++)
    if (t.elem(a) < 4) eq = eq && gl::canonical(cur[a]) == gl::canonical(cap[4 * index + t.elem(a)]);
  const bool all_eq = t.all(eq);
  if (have && t.elem(0) == 0) ok[q] = in_cap && all_eq;
}

}  // namespace pmt
