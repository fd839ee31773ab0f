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
#include "poseidon_quad.cuh"

namespace pmt {

using poseidon::WIDTH;

// the production permutation (see DESIGN.md "Permutation variants" for the measurements behind this choice).
// PMT_PERM selects the form for A/B runs (tools/ab_level.cu): 0 = permute_fast (sparse partial rounds), 1 = permute_fused,
// 2 = permute_rounds (30 x S-box + DFMA MDS), 3 = permute_paired (partial rounds in pairs; round-1 production until the
// frequency form), 4 = permute_paired_freq (3 with the MDS layers as frequency-domain convolutions, poseidon_freq.cuh;
// production: 1.55 against 1.29 G permutations/s, profiles/ab_freq_r1.jsonl).
#ifndef PMT_PERM
#define PMT_PERM 4
#endif
#ifndef PMT_SBOX_FMA_MASK
#define PMT_SBOX_FMA_MASK 0
#endif
#ifndef PMT_PART_FMA_MASK
#define PMT_PART_FMA_MASK 0
#endif
#ifndef PMT_MULADD_ALU
#define PMT_MULADD_ALU 1
#endif
#ifndef PMT_DOT_ALU
#define PMT_DOT_ALU 0
#endif
#ifndef PMT_CVT_I2F
#define PMT_CVT_I2F 1   // I2F.F64.U32 (conversion pipe) instead of the 2^52 magic-number subtraction (fma pipe)
#endif
#ifndef PMT_COMBINE_ALU
#define PMT_COMBINE_ALU 1
#endif
#ifndef PMT_COLUMN
#define PMT_COLUMN 0
#endif
#ifndef PMT_PIPE
#define PMT_PIPE 0
#endif
#ifndef PMT_PPIPE
#define PMT_PPIPE 0
#endif
#ifndef PMT_FQ_COMBINE
#define PMT_FQ_COMBINE 0   // recombination of the fp64 sums: 0 = ALU only, 1 = IMAD.WIDE folds, 2 = IMAD.WIDE in full layers only, 3 = in pairs only
#endif
#ifndef PMT_FQ_SPLIT
#define PMT_FQ_SPLIT 0   // 1: fence the high halves behind the low halves (measured slower: 1.51 against 1.55)
#endif
#ifndef PMT_COMPRESS_SPECIALISED
#define PMT_COMPRESS_SPECIALISED 0
#endif
template <bool CAP_ZERO, bool OUT4>
__device__ __forceinline__ void permute_impl(uint64_t (&s)[WIDTH]) {
#if PMT_PERM == 0
  poseidon::permute_fast<true, true, 2, CAP_ZERO, OUT4>(s);
#elif PMT_PERM == 3
  poseidon::permute_paired<PMT_SBOX_FMA_MASK, PMT_PART_FMA_MASK, PMT_COLUMN != 0, PMT_CVT_I2F != 0, PMT_COMBINE_ALU != 0, CAP_ZERO,
                           OUT4>(s);
#elif PMT_PERM == 4
  poseidon::permute_paired_freq<PMT_SBOX_FMA_MASK, PMT_PART_FMA_MASK, PMT_CVT_I2F != 0, CAP_ZERO, OUT4, PMT_FQ_SPLIT, PMT_FQ_COMBINE>(s);
#elif PMT_PERM == 2
  poseidon::permute_rounds<PMT_SBOX_FMA_MASK, PMT_PART_FMA_MASK, PMT_COLUMN != 0, PMT_CVT_I2F != 0, PMT_COMBINE_ALU != 0, CAP_ZERO,
                           OUT4>(s);
#else
  poseidon::permute_fused<PMT_SBOX_FMA_MASK, PMT_PART_FMA_MASK, PMT_MULADD_ALU != 0, PMT_DOT_ALU != 0, PMT_CVT_I2F != 0,
                          PMT_COMBINE_ALU != 0, CAP_ZERO, OUT4, PMT_PIPE, PMT_PPIPE>(s);
#endif
}
__device__ __forceinline__ void permute(uint64_t (&s)[WIDTH]) { permute_impl<false, false>(s); }
// two_to_one: zero capacity lanes on entry, only the digest lanes are read afterwards
__device__ __forceinline__ void permute_compress(uint64_t (&s)[WIDTH]) {
  permute_impl<PMT_COMPRESS_SPECIALISED != 0, PMT_COMPRESS_SPECIALISED != 0>(s);
}

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
  permute_compress(s);
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

// one level: digest(l, k) = two_to_one(children) for k in [k0, k0 + count)
// PMT_QUAD = 0 (production): one node per thread (poseidon.cuh permute_paired).
// PMT_QUAD = 1: 32 nodes per warp in the quad layout of poseidon_quad.cuh, MDS layers on the fp64 tensor pipe (DMMA).
// Thread (q, j) loads / stores element j of the digests of nodes 8 mb + q: the four threads of a quad cover one 32-byte
// digest.  Bit-exact and 30 % fewer instructions, but not faster on B200 (1.26 vs 1.29 G permutations/s):
// both forms are bound by the S-boxes' integer work, which the MDS engine does not change.
#ifndef PMT_QUAD
#define PMT_QUAD 0
#endif
#ifndef PMT_QUAD_SBOX_FMA_MASK
#define PMT_QUAD_SBOX_FMA_MASK 0
#endif
#ifndef PMT_QUAD_PART_FMA_MASK
#define PMT_QUAD_PART_FMA_MASK 0
#endif
#ifndef PMT_QUAD_COMBINE_ALU
#define PMT_QUAD_COMBINE_ALU 1
#endif
__device__ __forceinline__ void permute_quad(uint64_t (&e)[4][3], const poseidon::QuadTables& T, const poseidon::QuadFrags& f,
                                             unsigned lane) {
  poseidon::permute_quad<PMT_QUAD_SBOX_FMA_MASK, PMT_QUAD_PART_FMA_MASK, PMT_QUAD_COMBINE_ALU != 0>(e, T, f, lane);
}

template <class Layout>
__global__ void __launch_bounds__(BLOCK, PMT_MINB) k_level(Layout lay, int l, size_t k0, size_t count) {
#if PMT_QUAD
  __shared__ poseidon::QuadTables T;
  poseidon::quad_stage_tables(T);
  const unsigned lane = threadIdx.x & 31, q = lane >> 2, j = lane & 3;
  poseidon::QuadFrags f;
  poseidon::quad_load_frags(f, q, j);
#if PMT_QUAD_FRAGS_SMEM
  poseidon::quad_publish_frags(T, f, lane);
#endif
  for (size_t base = (size_t)blockIdx.x * BLOCK; base < count; base += (size_t)gridDim.x * BLOCK) {
    const size_t wbase = base + (threadIdx.x & ~31u);
    if (wbase >= count) continue;                       // warp-uniform
    uint64_t e[4][3];
#pragma unroll
    for (int mb = 0; mb < 4; mb++) {
      size_t i = wbase + 8 * mb + q;
      if (i >= count) i = count - 1;                    // ragged tail: recompute the last node, store nothing
      const uint64_t *a, *b;
      lay.children(l, k0 + i, a, b);
      e[mb][0] = a[j]; e[mb][1] = b[j]; e[mb][2] = 0;
    }
    permute_quad(e, T, f, lane);
#pragma unroll
    for (int mb = 0; mb < 4; mb++) {
      const size_t i = wbase + 8 * mb + q;
      if (i < count) lay.at(l, k0 + i)[j] = gl::canonical(e[mb][0]);
    }
  }
#else
  for (size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x; i < count; i += (size_t)gridDim.x * BLOCK) {
    const uint64_t *a, *b;
    lay.children(l, k0 + i, a, b);
    store_digest(lay.at(l, k0 + i), two_to_one(load_digest(a), load_digest(b)));
  }
#endif
}

constexpr int TOP_BLOCK = 256;  // 256 x ~100 registers fits one SM's register file without spilling the state

// fused upper levels, ONE block: levels l0 .. l1 (inclusive); level l has count0 >> (l - l0) nodes starting at node 0.
// Children written by this block in the previous iteration are read back through L2 after a block barrier.
template <class Layout>
__global__ void __launch_bounds__(TOP_BLOCK) k_top(Layout lay, int l0, int l1, size_t count0) {
  size_t count = count0;
  for (int l = l0; l <= l1; l++, count >>= 1) {
    for (size_t i = threadIdx.x; i < count; i += blockDim.x) {
      const uint64_t *a, *b;
      lay.children(l, i, a, b);
      store_digest(lay.at(l, i), two_to_one(load_digest(a), load_digest(b)));
    }
    __threadfence_block();
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// cooperative (16 lanes per permutation) kernels for levels too small to fill the GPU with one thread per node
// ---------------------------------------------------------------------------------------------------------------
constexpr int COOP_BLOCK = 256;          // 16 permutations in flight per block
constexpr int COOP_GROUPS = COOP_BLOCK / 16;

__device__ __forceinline__ void coop_stage_constants(uint64_t* rc_smem) {
  for (int i = threadIdx.x; i < WIDTH * (PMT_ROUNDS + 1); i += blockDim.x) rc_smem[i] = PMT_RC[i];
  __syncthreads();
}

// one two_to_one by a 16-lane group; `active` = this group has a node (inactive groups still run the shuffles)
template <class Layout>
__device__ __forceinline__ void coop_node(const Layout& lay, int l, size_t k, bool active, const uint64_t* rc_smem,
                                          unsigned g, unsigned base_lane) {
  uint64_t v = 0;
  if (active && g < 8) {
    const uint64_t *a, *b;
    lay.children(l, k, a, b);
    v = g < 4 ? a[g] : b[g - 4];
  }
  v = poseidon::permute_coop(v, rc_smem, g, base_lane);
  if (active && g < 4) lay.at(l, k)[g] = gl::canonical(v);
}

template <class Layout>
__global__ void __launch_bounds__(COOP_BLOCK) k_level_coop(Layout lay, int l, size_t k0, size_t count) {
  __shared__ uint64_t rc_smem[WIDTH * (PMT_ROUNDS + 1)];
  coop_stage_constants(rc_smem);
  const unsigned g = threadIdx.x & 15, base_lane = threadIdx.x & 16;
  const size_t group = (size_t)blockIdx.x * COOP_GROUPS + (threadIdx.x >> 4);
  const size_t stride = (size_t)gridDim.x * COOP_GROUPS;
  // trip count is uniform across the block so every warp executes the same shuffles
  for (size_t base = 0; base < count; base += stride) {
    const size_t i = base + group;
    coop_node(lay, l, k0 + i, i < count, rc_smem, g, base_lane);
  }
}

// fused upper levels in ONE block of 256 threads (16 groups): levels l0 .. l1, level l has count0 >> (l - l0) nodes.
// Warps whose two groups both have no node skip the permutation (shuffles never cross a warp).
template <class Layout>
__global__ void __launch_bounds__(COOP_BLOCK) k_top_coop(Layout lay, int l0, int l1, size_t count0) {
  __shared__ uint64_t rc_smem[WIDTH * (PMT_ROUNDS + 1)];
  coop_stage_constants(rc_smem);
  const unsigned g = threadIdx.x & 15, base_lane = threadIdx.x & 16;
  const size_t group = threadIdx.x >> 4, groups = blockDim.x >> 4;
  size_t count = count0;
  for (int l = l0; l <= l1; l++, count >>= 1) {
    for (size_t base = 0; base < count; base += groups) {
      const size_t first_of_warp = base + (group & ~(size_t)1);
      if (first_of_warp < count) coop_node(lay, l, base + group, base + group < count, rc_smem, g, base_lane);
    }
    __threadfence_block();
    __syncthreads();
  }
}

// fused subtree blocks for the latency-bound middle of a tree (levels of <= 2^13 nodes): block b owns the 16 nodes
// k0 + [16 b, 16 b + 16) of level l0 (k0 a multiple of 16) and every ancestor of theirs up to `levels` levels (16, 8, 4, 2, 1 nodes), one
// 16-lane group per node, shuffles inside the permutation, a block barrier between levels.  A block only ever reads
// children it wrote itself (through L1/L2; every level has to be stored anyway, proofs need it), so the blocks are
// independent and one launch replaces up to five.  Warps whose two groups both have no node skip the permutation.
template <class Layout>
__global__ void __launch_bounds__(COOP_BLOCK) k_subtree_coop(Layout lay, int l0, int levels, size_t k0) {
  __shared__ uint64_t rc_smem[WIDTH * (PMT_ROUNDS + 1)];
  coop_stage_constants(rc_smem);
  const unsigned g = threadIdx.x & 15, base_lane = threadIdx.x & 16;
  const size_t group = threadIdx.x >> 4;
  size_t per = COOP_GROUPS;                                  // nodes of this block at the current level
  for (int j = 0; j < levels; j++, per >>= 1) {
    if ((group & ~(size_t)1) < per) coop_node(lay, l0 + j, (k0 >> j) + (size_t)blockIdx.x * per + group, group < per, rc_smem, g, base_lane);
    __threadfence_block();
    __syncthreads();
  }
}

// multi-GPU finish: the levels above n_roots gathered subtree roots, ONE block, 16 lanes per permutation (the levels are
// sequential and tiny: G - 1 permutations for G ranks).  out is level-major: n_roots/2, n_roots/4, ..., n_cap digests.
// blockIdx.x = which set of roots (a batch of independent finishes: the rounds of a sharded MMR); sets are n_roots digests
// apart in `roots` and n_roots - n_cap digests apart in `out`.
__global__ void __launch_bounds__(COOP_BLOCK) k_top_roots_coop(const uint64_t* __restrict__ roots, size_t n_roots, size_t n_cap,
                                                               uint64_t* __restrict__ out) {
  __shared__ uint64_t rc_smem[WIDTH * (PMT_ROUNDS + 1)];
  coop_stage_constants(rc_smem);
  roots += 4 * n_roots * blockIdx.x;
  out += 4 * (n_roots - n_cap) * blockIdx.x;
  const unsigned g = threadIdx.x & 15, base_lane = threadIdx.x & 16;
  const size_t group = threadIdx.x >> 4, groups = blockDim.x >> 4;
  const uint64_t* cur = roots;
  uint64_t* dst = out;
  for (size_t m = n_roots / 2; m >= n_cap && m >= 1; m >>= 1) {
    for (size_t base = 0; base < m; base += groups) {
      const size_t first_of_warp = base + (group & ~(size_t)1);
      if (first_of_warp < m) {                       // warp-uniform: both groups of a warp run the shuffles
        const size_t k = base + group;
        const bool active = k < m;
        uint64_t v = 0;
        if (active && g < 8) v = cur[8 * k + g];     // children 2k and 2k + 1 are adjacent
        v = poseidon::permute_coop(v, rc_smem, g, base_lane);
        if (active && g < 4) dst[4 * k + g] = gl::canonical(v);
      }
    }
    __threadfence_block();
    __syncthreads();
    cur = dst;
    dst += 4 * m;
    if (m == 1) break;
  }
}

// hash_or_noop of ONE row by one 16-lane group (bagging the peaks): the sponge's permutations are sequential, so the
// cooperative form cuts the latency by ~10x
__global__ void __launch_bounds__(32) k_hash_one_coop(const uint64_t* __restrict__ felts, size_t w, uint64_t* __restrict__ out) {
  __shared__ uint64_t rc_smem[WIDTH * (PMT_ROUNDS + 1)];
  coop_stage_constants(rc_smem);
  const unsigned g = threadIdx.x & 15, base_lane = threadIdx.x & 16;
  if (w <= 4) {
    if (threadIdx.x < 4) out[threadIdx.x] = threadIdx.x < w ? gl::canonical(felts[threadIdx.x]) : 0ull;
    return;
  }
  uint64_t v = 0;
  for (size_t off = 0; off < w; off += 8) {
    if (g < 8 && off + g < w) v = felts[off + g];      // overwrite mode: lanes past the chunk keep their state
    v = poseidon::permute_coop(v, rc_smem, g, base_lane);
  }
  if (threadIdx.x < 4) out[threadIdx.x] = gl::canonical(v);
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
// proofs: pure gathers, one thread per (proof, level)
// ---------------------------------------------------------------------------------------------------------------
// simple_merkle_tree.rs:55-74 get_merkle_proof: level_i[idx_i ^ 1], i = 0 .. log2(n) - 1
__global__ void k_simple_prove(LevelMajor lay, const uint64_t* __restrict__ idx, size_t n_idx, uint64_t* __restrict__ out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int depth = lay.top;
  if (t >= n_idx * (size_t)depth) return;
  const size_t q = t / depth; const int l = (int)(t % depth);
  const size_t k = (idx[q] >> l) ^ 1;
  store_digest(out + 4 * t, load_digest(lay.at(l, k)));
}

// [UPSTREAM hash/merkle_tree.rs MerkleTree::prove]
__global__ void k_plonky2_prove(Plonky2 lay, const uint64_t* __restrict__ idx, size_t n_idx, uint64_t* __restrict__ out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int depth = lay.sub_levels;
  if (t >= n_idx * (size_t)depth) return;
  const size_t q = t / depth; const int l = (int)(t % depth);
  const size_t k = (idx[q] >> l) ^ 1;
  store_digest(out + 4 * t, load_digest(lay.at(l, k)));
}

// merkle_mountain_ranges.rs:147-176 get_subtree_proof_elm, closed form: the leaf's mountain has height H = the bit of
// n_leaves that covers it; entry j = (node (j, (i >> j) ^ 1), sibling_on_left = bit j of i)
__global__ void k_mmr_prove(Mmr lay, size_t n_leaves, const uint64_t* __restrict__ idx, size_t n_idx,
                            uint64_t* __restrict__ sib_out, uint8_t* __restrict__ left_out, uint32_t* __restrict__ len_out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_idx * 32) return;
  const size_t q = t / 32; const int j = (int)(t % 32);
  const size_t i = idx[q];
  // mountains from the largest: the leaf is in the first mountain whose end exceeds i
  int H = 0; size_t base = 0;
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

// merkle_mountain_ranges.rs:122-127 bagging_the_peaks = hash_or_noop(flattened peaks); one thread (<= 32 permutations)
__global__ void k_hash_one(const uint64_t* __restrict__ felts, size_t w, uint64_t* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) store_digest(out, hash_or_noop(felts, w));
}

// ---------------------------------------------------------------------------------------------------------------
// verification: one thread folds one path
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

// simple_merkle_tree.rs:91-109 verify_merkle_proof and [UPSTREAM hash/merkle_proofs.rs verify_merkle_proof_to_cap]:
// fold by index parity, compare with cap[index >> path_len] (cap_height 0 + width 1 = the simple tree's root)
__global__ void __launch_bounds__(BLOCK, PMT_MINB) k_verify_to_cap(const uint64_t* __restrict__ rows, size_t w,
                                                         const uint64_t* __restrict__ idx, size_t n_idx,
                                                         const uint64_t* __restrict__ cap, uint32_t cap_height,
                                                         const uint64_t* __restrict__ proofs, size_t path_len,
                                                         uint8_t* __restrict__ ok) {
  const size_t q = (size_t)blockIdx.x * BLOCK + threadIdx.x;
  if (q >= n_idx) return;
  size_t index = idx[q];
  Digest cur = hash_or_noop(rows + q * w, w);
  for (size_t j = 0; j < path_len; j++) {
    const Digest sib = load_digest(proofs + 4 * (q * path_len + j));
    cur = (index & 1) ? two_to_one(sib, cur) : two_to_one(cur, sib);
    index >>= 1;
  }
  ok[q] = (index < ((size_t)1 << cap_height)) && digest_eq(cur, load_digest_canonical(cap + 4 * index));
}

// merkle_mountain_ranges.rs:232-252 MMR_proof::verify: fold by sibling_on_left, membership in peaks (else the
// reference panics: status -1), re-bag, compare with root
__global__ void __launch_bounds__(BLOCK, PMT_MINB) k_mmr_verify(const uint64_t* __restrict__ leaves, size_t n_idx,
                                                      const uint64_t* __restrict__ sib, const uint8_t* __restrict__ left,
                                                      const uint32_t* __restrict__ len, const uint64_t* __restrict__ peaks,
                                                      uint32_t n_peaks, const uint64_t* __restrict__ bagged,
                                                      const uint64_t* __restrict__ root, int8_t* __restrict__ status) {
  const size_t q = (size_t)blockIdx.x * BLOCK + threadIdx.x;
  if (q >= n_idx) return;
  Digest cur = hash_or_noop(leaves + q, 1);
  const uint32_t L = len[q];
  for (uint32_t j = 0; j < L; j++) {
    const Digest s = load_digest(sib + 4 * (q * 32 + j));
    cur = left[q * 32 + j] ? two_to_one(s, cur) : two_to_one(cur, s);
  }
  bool found = false;
  for (uint32_t k = 0; k < n_peaks; k++) found |= digest_eq(cur, load_digest_canonical(peaks + 4 * k));
  if (!found) { status[q] = -1; return; }
  // the re-bagged root is identical for every proof of the batch: computed once by k_hash_one into `bagged`
  status[q] = digest_eq(load_digest(bagged), load_digest_canonical(root)) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------------
// cooperative verification (batches too small to fill the GPU with one thread per proof): 16 lanes fold one path, so a
// proof of length L costs L x 6.6 us instead of L x 38 us.  Lanes 0..3 of a group carry the running digest.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool group_all(bool pred, unsigned base_lane) {
  const unsigned b = __ballot_sync(0xffffffffu, pred), mask = 0xFFFFu << base_lane;
  return (b & mask) == mask;
}
// one fold step: v <- two_to_one(sibling, v) if on_left else two_to_one(v, sibling); `sib` points at the sibling digest
__device__ __forceinline__ uint64_t coop_fold(uint64_t v, const uint64_t* __restrict__ sib, bool on_left, bool active,
                                              const uint64_t* rc_smem, unsigned g, unsigned base_lane) {
  const uint64_t moved = gl::pack(__shfl_sync(0xffffffffu, gl::lo32(v), base_lane + ((g - 4) & 15)),
                                  __shfl_sync(0xffffffffu, gl::hi32(v), base_lane + ((g - 4) & 15)));   // lane g <- lane g - 4
  uint64_t st = 0;
  if (on_left) { if (g < 4) st = active ? sib[g] : 0; else if (g < 8) st = moved; }
  else         { if (g < 4) st = v; else if (g < 8) st = active ? sib[g - 4] : 0; }
  return poseidon::permute_coop(st, rc_smem, g, base_lane);
}

// merkle_mountain_ranges.rs:232-252, same contract as k_mmr_verify
__global__ void __launch_bounds__(COOP_BLOCK) k_mmr_verify_coop(const uint64_t* __restrict__ leaves, size_t n_idx,
                                                                const uint64_t* __restrict__ sib, const uint8_t* __restrict__ left,
                                                                const uint32_t* __restrict__ len, const uint64_t* __restrict__ peaks,
                                                                uint32_t n_peaks, const uint64_t* __restrict__ bagged,
                                                                const uint64_t* __restrict__ root, int8_t* __restrict__ status) {
  __shared__ uint64_t rc_smem[WIDTH * (PMT_ROUNDS + 1)];
  coop_stage_constants(rc_smem);
  const unsigned g = threadIdx.x & 15, base_lane = threadIdx.x & 16;
  const size_t q = (size_t)blockIdx.x * COOP_GROUPS + (threadIdx.x >> 4);
  const bool have = q < n_idx;
  const uint32_t L = have ? len[q] : 0;
  const uint32_t Lw = max(L, __shfl_xor_sync(0xffffffffu, L, 16));     // both groups of the warp run the same trip count
  uint64_t v = (have && g == 0) ? gl::canonical(leaves[q]) : 0ull;     // hash_or_noop(&[leaf]) = [leaf, 0, 0, 0]
  for (uint32_t j = 0; j < Lw; j++) {
    const bool act = j < L;
    const bool on_left = act && left[q * 32 + j] != 0;
    const uint64_t p = coop_fold(v, sib + 4 * (q * 32 + j), on_left, act, rc_smem, g, base_lane);
    v = act ? p : v;
  }
  v = gl::canonical(v);
  bool found = false;
  for (uint32_t k = 0; k < n_peaks; k++) found |= group_all(g >= 4 || v == gl::canonical(peaks[4 * k + g]), base_lane);
  if (have && g == 0) {
    if (!found) status[q] = -1;
    else status[q] = digest_eq(load_digest(bagged), load_digest_canonical(root)) ? 1 : 0;
  }
}

// simple_merkle_tree.rs:91-109 / [UPSTREAM hash/merkle_proofs.rs verify_merkle_proof_to_cap], same contract as k_verify_to_cap
__global__ void __launch_bounds__(COOP_BLOCK) k_verify_to_cap_coop(const uint64_t* __restrict__ rows, size_t w,
                                                                   const uint64_t* __restrict__ idx, size_t n_idx,
                                                                   const uint64_t* __restrict__ cap, uint32_t cap_height,
                                                                   const uint64_t* __restrict__ proofs, size_t path_len,
                                                                   uint8_t* __restrict__ ok) {
  __shared__ uint64_t rc_smem[WIDTH * (PMT_ROUNDS + 1)];
  coop_stage_constants(rc_smem);
  const unsigned g = threadIdx.x & 15, base_lane = threadIdx.x & 16;
  const size_t q = (size_t)blockIdx.x * COOP_GROUPS + (threadIdx.x >> 4);
  const bool have = q < n_idx;
  const uint64_t* row = rows + (have ? q : 0) * w;
  uint64_t v = 0;
  if (w <= 4) {
    if (have && g < w) v = gl::canonical(row[g]);
  } else {
    for (size_t off = 0; off < w; off += 8) {            // overwrite-mode sponge; w is uniform, so is the trip count
      if (have && g < 8 && off + g < w) v = row[off + g];
      v = poseidon::permute_coop(v, rc_smem, g, base_lane);
    }
  }
  size_t index = have ? idx[q] : 0;
  for (size_t j = 0; j < path_len; j++) {
    v = coop_fold(v, proofs + 4 * ((have ? q : 0) * path_len + j), (index & 1) != 0, have, rc_smem, g, base_lane);
    index >>= 1;
  }
  v = gl::canonical(v);
  const bool in_cap = index < ((size_t)1 << cap_height);
  const bool eq = group_all(g >= 4 || (in_cap && have && v == gl::canonical(cap[4 * index + g])), base_lane);
  if (have && g == 0) ok[q] = in_cap && eq;
}

}  // namespace pmt
