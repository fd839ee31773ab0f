// poseidon_coop.cuh -- the same Poseidon permutation with SEVERAL threads per state, for the latency-bound part of the path.
//
// Same function as poseidon.cuh's permute ([UPSTREAM plonky2 hash/poseidon.rs Poseidon::poseidon] under
// PoseidonHash::{two_to_one, hash_or_noop} at /root/reference/src/simple_merkle_tree/simple_merkle_tree.rs:23,33,45 and
// /root/reference/src/mmr/merkle_mountain_ranges.rs:96,111,125,238-249), other mappings.
//
// Why.  A tree level with few nodes, a proof path, a peak bag are chains of DEPENDENT permutations: what counts is the
// latency of one permutation, not the throughput of many.  One state per thread takes a lone warp 33 us (15.5 k
// instructions).  Two forms split a state over threads (measured on B200, tools/perm_bench.cu lat, profiles/latency_r2*.jsonl):
//
//   Quad (4 threads x 3 elements, 8 states per warp) -- the THROUGHPUT form of the cooperative kernels (levels of 2^5 .. 2^13
//     nodes, proof batches): thread j holds elements j, 4 + j, 8 + j.  Three independent S-boxes per thread in a full round.
//     The MDS layer runs on the fp64 pipe like the thread-per-state form: the owner of an element converts its 32-bit halves
//     to doubles (I2F) and publishes them in a per-warp shared-memory exchange buffer, one __syncwarp, and every thread
//     accumulates ITS three output rows over the 12 published elements, read in ROTATED order (slot i = element (j + i) mod
//     12; elements 0..2 are stored twice so the rotation needs no address arithmetic) -- with that order the coefficient of
//     slot i in row j + 4a is CIRC[(i - 4a) mod 12] for every thread, a uniform operand: 72 DFMAs per thread and layer.
//     Accumulators start at 2^52 + round constant, so the mantissa is the integer (combine_magic).  The buffer is double-
//     buffered: one __syncwarp per round; 15 slots per state (odd) keep writes and rotated reads free of bank conflicts.
//     870 warp instructions per state (round 1's 16-lane form: 2 120), 0.53 G permutations/s, 11 us per permutation.
//   Wide (16 lanes x 1 element, 12 used, 2 states per warp) -- the LATENCY form (levels of <= 16 nodes per block, the top
//     of a tree, bagging, small proof batches): everything stays in integer registers, because in a lone warp the fp64 /
//     shared-memory layer costs ~300 cycles of dependent latency per round (I2F, STS -> LDS, DFMA chains) against ~150 for
//     22 warp shuffles feeding four independent IMAD.WIDE accumulation chains (column sums < 2^42: no carries) and one
//     reduction.  An fp64 Wide form was built and measured: 9.4 us per permutation against 6.2 us for round 1's integer form.
//
// Partial rounds, both forms: the linear part of lanes 1..11 does not depend on the S-box of lane 0.  The rows are
// accumulated over the other eleven elements (the owner of element 0 contributes zero) while the S-box chain runs -- the
// S-box is written after the exchange so that both land in one basic block and ptxas interleaves them -- then x = sbox(s0)
// is broadcast with two shuffles and enters with the per-thread coefficients M[row][0]: the round's critical path is the
// S-box plus a dozen instructions instead of S-box plus layer.
//
// Exactness (Quad): products of a 32-bit half with a coefficient < 64 summed over 12 lanes plus the diagonal term and a
// 32-bit constant half stay below 2^42 -- exact in the 53-bit mantissa (the same bound as the matrix-form fp64 layer of
// round 1); all values are integers, so the order of the additions does not matter.
#pragma once
#include "poseidon.cuh"

namespace poseidon {
namespace coop {

// Exchange buffer of one warp: 15 sixteen-byte slots per state (elements 0..11, then copies of 0..2).  A 128-bit shared
// access is served a QUARTER-warp (8 threads = two quads) at a time, each 16-byte slot covering a group of 4 banks, so
// the two quads of a quarter must touch disjoint slots mod 8: the odd quad of a pair sits 20 slots (= 4 mod 8) after the
// even one, pairs are 35 slots apart.  (The first version used a flat stride of 15: two-way conflicts on every access,
// 1 900 excess wavefronts per permutation in profiles/k_perm_form_r2_summary.txt.)
constexpr int QUAD_PAIR_STRIDE = 35, QUAD_ODD_OFFSET = 20;
constexpr int WARP_SLOTS = 4 * QUAD_PAIR_STRIDE;     // one exchange buffer of one warp
__device__ __forceinline__ int quad_slot_base(unsigned quad_in_warp) {
  return (int)(quad_in_warp >> 1) * QUAD_PAIR_STRIDE + (int)(quad_in_warp & 1) * QUAD_ODD_OFFSET;
}

template <int WARPS>
struct alignas(16) Shared {
  double2 xch[WARPS][2][WARP_SLOTS];       // exchange buffers (lo half, hi half of an element as doubles)
  double rc_dm[2 * WIDTH * PMT_ROUNDS];    // PMT_RC_DM: 2^52 + halves of the constants the layer of round r adds (Quad)
  uint64_t rc[WIDTH * (PMT_ROUNDS + 1)];   // PMT_RC: row 0 = the first constant layer, row r + 1 = what round r's layer adds (Wide)
};

// The round constants in GLOBAL memory, for staging: per-lane indices into constant memory serialise (32 replays per warp
// load, and the 8.7 KB of tables miss the constant cache): staging them from the constant bank cost ~8 us at the head of
// every cooperative launch.  From global memory (L2-resident after the first launch) it is one coalesced pass.  Filled
// once per device by pmt_init (k_coop_tables_init).
static __device__ double RC_DM_G[2 * WIDTH * PMT_ROUNDS];
static __device__ uint64_t RC_G[WIDTH * (PMT_ROUNDS + 1)];
static __global__ void k_coop_tables_init() {
  for (int i = threadIdx.x; i < 2 * WIDTH * PMT_ROUNDS; i += blockDim.x) RC_DM_G[i] = PMT_RC_DM[i];
  for (int i = threadIdx.x; i < WIDTH * (PMT_ROUNDS + 1); i += blockDim.x) RC_G[i] = PMT_RC[i];
}

// all threads of the block; ends with a block barrier
template <int WARPS>
__device__ __forceinline__ void stage(Shared<WARPS>& sh) {
  for (int i = threadIdx.x; i < 2 * WIDTH * PMT_ROUNDS; i += blockDim.x) sh.rc_dm[i] = RC_DM_G[i];
  for (int i = threadIdx.x; i < WIDTH * (PMT_ROUNDS + 1); i += blockDim.x) sh.rc[i] = RC_G[i];
  __syncthreads();
}

// the 32-bit halves of x as doubles.  2^52 magic-number conversion (pair the word with 0x43300000, subtract 2^52): one
// DADD per half on the fp64 pipe.  The thread-per-state kernel uses I2F on the otherwise idle conversion pipe (better
// THROUGHPUT), but I2F + its scoreboard wait is 35 cycles of dependent latency against 9 for the DADD, and here latency is
// what counts (tools/lat_bench.cu).
__device__ __forceinline__ double2 halves(uint64_t x) {
  const double MAGIC = 4503599627370496.0;
  return make_double2(__hiloint2double(0x43300000, (int)gl::lo32(x)) - MAGIC, __hiloint2double(0x43300000, (int)gl::hi32(x)) - MAGIC);
}

// ---------------------------------------------------------------------------------------------------------------
// Quad: thread j of 4 adjacent lanes holds e[a] = element j + 4a
// ---------------------------------------------------------------------------------------------------------------
struct Quad {
  static constexpr int LANES = 4, ELEMS = 3;
  unsigned j;            // position in the quad
  unsigned lane0;        // warp lane of the quad's thread 0
  double2* slot;         // &xch[warp][0][quad_slot_base(quad) + j]: writes at +0 (+12 for j < 3), +4, +8; reads at +0 .. +11
  const double* kdm;     // rc_dm + 2 j: the constants of element j + 4a in round r are kdm[24 r + 8 a + {0, 1}]
  double c00;            // coefficient of slot 0 in output row j: C[0] + DIAG[0] = 25 on thread 0, C[0] = 17 elsewhere
  double cx[3];          // M[j + 4a][0]: coefficient of element 0 in this thread's three output rows

  template <int WARPS>
  __device__ __forceinline__ static Quad make(Shared<WARPS>& sh) {
    Quad t;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    t.j = lane & 3;
    t.lane0 = lane & ~3u;
    t.slot = &sh.xch[warp][0][quad_slot_base(lane >> 2) + t.j];
    t.kdm = sh.rc_dm + 2 * t.j;
    t.c00 = t.j == 0 ? 25.0 : 17.0;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const unsigned r = t.j + 4 * a;
      t.cx[a] = PMT_MDS_CIRC_D[r == 0 ? 12 : 12 - r];
    }
    return t;
  }
  // index of the state element held in e[a]
  __device__ __forceinline__ unsigned elem(int a) const { return j + 4 * a; }
  // position of this thread's state among the states of the block
  __device__ __forceinline__ unsigned state() const { return threadIdx.x >> 2; }
  __device__ __forceinline__ bool all(bool pred) const {          // pred on every thread of the quad
    const unsigned b = __ballot_sync(0xffffffffu, pred), mask = 0xFu << lane0;
    return (b & mask) == mask;
  }

  // one MDS layer of this thread's three rows over the 12 published slots; k = the layer's constants
  __device__ __forceinline__ void rows(const double2* __restrict__ rd, const double* __restrict__ k, double (&L)[3], double (&H)[3]) const {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const double2 c = *reinterpret_cast<const double2*>(k + 8 * a);
      L[a] = c.x; H[a] = c.y;
    }
#pragma unroll
    for (int i = 0; i < WIDTH; i++) {
      const double2 v = rd[i];
#pragma unroll
      for (int a = 0; a < 3; a++) {
        const double c = (i == 0 && a == 0) ? c00 : PMT_MDS_CIRC_D[(i - 4 * a + WIDTH) % WIDTH];
        L[a] = fma(v.x, c, L[a]);
        H[a] = fma(v.y, c, H[a]);
      }
    }
  }

  // The permutation.  Every lane of the warp must call it (quads without work pass zeros).  Outputs NOT canonicalised.
  template <int WARPS>
  __device__ __forceinline__ void permute(uint64_t (&e)[3], const Shared<WARPS>& sh) const {
#pragma unroll
    for (int a = 0; a < 3; a++) e[a] = gl::add_canonical(e[a], sh.rc[j + 4 * a]);
    double2* s = slot;
    const double* k = kdm;
    int toggle = WARP_SLOTS;   // s alternates between the two buffers: +WARP_SLOTS, -WARP_SLOTS, ...
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
#pragma unroll 1
      for (int q = 0; q < PMT_FULL_HALF; q++) {
#pragma unroll
        for (int a = 0; a < 3; a++) e[a] = gl::pow7(e[a]);
        const double2 d0 = halves(e[0]);
        s[0] = d0;
        if (j < 3) s[WIDTH] = d0;
        s[4] = halves(e[1]);
        s[8] = halves(e[2]);
        __syncwarp();
        double L[3], H[3];
        rows(s, k, L, H);
#pragma unroll
        for (int a = 0; a < 3; a++) e[a] = combine_magic(L[a], H[a]);
        s += toggle; toggle = -toggle; k += 2 * WIDTH;
      }
      if (half == 0) {
#pragma unroll 1
        for (int p = 0; p < PMT_PARTIAL; p++) {
          const double2 d0 = j == 0 ? make_double2(0.0, 0.0) : halves(e[0]);
          s[0] = d0;
          if (j < 3) s[WIDTH] = d0;
          s[4] = halves(e[1]);
          s[8] = halves(e[2]);
          __syncwarp();
          double L[3], H[3];
          rows(s, k, L, H);                             // independent of x: overlaps the S-box chain below
          const uint64_t x = gl::pow7(e[0]);            // only thread 0's is used; the others' element j skips the S-box
          const double2 dx = halves(gl::pack(__shfl_sync(0xffffffffu, gl::lo32(x), lane0), __shfl_sync(0xffffffffu, gl::hi32(x), lane0)));
          const double xl = dx.x, xh = dx.y;
#pragma unroll
          for (int a = 0; a < 3; a++) {
            L[a] = fma(xl, cx[a], L[a]);
            H[a] = fma(xh, cx[a], H[a]);
            e[a] = combine_magic(L[a], H[a]);
          }
          s += toggle; toggle = -toggle; k += 2 * WIDTH;
        }
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Wide: lane g of 16 adjacent lanes holds e[0] = element g (g < 12; lanes 12..15 carry nothing).  The LATENCY form: what
// counts is the dependent chain of one round, so everything stays in integer registers -- no conversions, no shared-memory
// round trip, no fp64 chain (measured: the fp64 / shared-memory layer costs ~300 cycles of latency per round in a lone
// warp against ~150 for shuffles + IMAD.WIDE): the MDS row of a lane is 11 pairs of warp shuffles feeding four independent
// IMAD.WIDE accumulation chains (sums < 2^42: no carries), recombined with one reduction.
// ---------------------------------------------------------------------------------------------------------------
struct Wide {
  static constexpr int LANES = 16, ELEMS = 1;
  unsigned g;            // position in the group; g >= 12: idle (mirrors element 0, stores nothing)
  unsigned gg;           // g for working lanes, 0 for idle ones
  unsigned lane0;        // warp lane of the group's lane 0
  uint32_t c0;           // coefficient of the lane's own element in its row: C[0] + DIAG[0] = 25 on lane 0, C[0] = 17 elsewhere
  uint32_t cx;           // M[g][0]: coefficient of element 0 in this lane's row

  template <int WARPS>
  __device__ __forceinline__ static Wide make(Shared<WARPS>&) {
    Wide t;
    const unsigned lane = threadIdx.x & 31;
    t.g = lane & 15;
    t.gg = t.g < WIDTH ? t.g : 0;
    t.lane0 = lane & 16u;
    t.c0 = t.gg == 0 ? 25u : 17u;
    t.cx = PMT_MDS_CIRC32[t.gg == 0 ? 12 : 12 - t.gg];
    return t;
  }
  __device__ __forceinline__ unsigned elem(int) const { return g; }
  __device__ __forceinline__ unsigned state() const { return threadIdx.x >> 4; }
  __device__ __forceinline__ bool all(bool pred) const {
    const unsigned b = __ballot_sync(0xffffffffu, pred), mask = 0xFFFFu << lane0;
    return (b & mask) == mask;
  }

  // row g of the MDS layer over the state held one element per lane: out = k + sum_i s[(g + i) % 12] C[i] (+ 8 s[0] on lane 0).
  // SKIP0: element 0 does not take part (partial rounds: it enters later, after its S-box) -- lane 0 simply contributes zero
  // to its own term and to every shuffle that reads it.
  template <bool SKIP0>
  __device__ __forceinline__ void row(uint64_t v, uint64_t k, uint64_t& L, uint64_t& H) const {
    constexpr uint32_t CIRC[WIDTH] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    const bool mute = SKIP0 && gg == 0;
    const uint32_t lo = mute ? 0u : gl::lo32(v), hi = mute ? 0u : gl::hi32(v);
    uint64_t La = gl::mad_wide(lo, c0, k), Ha = (uint64_t)hi * c0, Lb = 0, Hb = 0;
#pragma unroll
    for (int i = 1; i < WIDTH; i++) {
      const unsigned idx = gg + i;
      const unsigned src = lane0 + (idx >= WIDTH ? idx - WIDTH : idx);
      const uint32_t slo = __shfl_sync(0xffffffffu, lo, src), shi = __shfl_sync(0xffffffffu, hi, src);
      if (i & 1) { Lb = gl::mad_wide(slo, CIRC[i], Lb); Hb = gl::mad_wide(shi, CIRC[i], Hb); }
      else       { La = gl::mad_wide(slo, CIRC[i], La); Ha = gl::mad_wide(shi, CIRC[i], Ha); }
    }
    L = La + Lb; H = Ha + Hb;
  }

  template <int WARPS>
  __device__ __forceinline__ void permute(uint64_t (&e)[1], const Shared<WARPS>& sh) const {
    uint64_t v = g < WIDTH ? gl::add_canonical(e[0], sh.rc[gg]) : 0ull;
    const uint64_t* k = sh.rc + WIDTH + gg;     // the constants the layer of round r adds: row r + 1 of the table
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
#pragma unroll 1
      for (int q = 0; q < PMT_FULL_HALF; q++) {
        v = gl::pow7(v);
        uint64_t L, H;
        row<false>(v, *k, L, H);                // constants < 2^64 - 2^48, column sums < 2^42: no overflow
        v = gl::combine_sums(L, H);
        k += WIDTH;
      }
      if (half == 0) {
#pragma unroll 1
        for (int p = 0; p < PMT_PARTIAL; p++) {
          uint64_t L, H;
          row<true>(v, *k, L, H);               // everything but element 0: independent of the S-box below, overlaps it
          const uint64_t x = gl::pow7(v);       // only lane 0's is used
          const uint32_t xl = __shfl_sync(0xffffffffu, gl::lo32(x), lane0), xh = __shfl_sync(0xffffffffu, gl::hi32(x), lane0);
          v = gl::combine_sums(gl::mad_wide(xl, cx, L), gl::mad_wide(xh, cx, H));
          k += WIDTH;
        }
      }
    }
    e[0] = v;
  }
};

}  // namespace coop
}  // namespace poseidon
