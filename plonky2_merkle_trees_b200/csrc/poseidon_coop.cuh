// poseidon_coop.cuh -- the same Poseidon permutation with SEVERAL threads per state, for the latency-bound part of the path.
//
// Same function as poseidon.cuh's permute ([UPSTREAM plonky2 hash/poseidon.rs Poseidon::poseidon] under
// PoseidonHash::{two_to_one, hash_or_noop} at /root/reference/src/simple_merkle_tree/simple_merkle_tree.rs:23,33,45 and
// /root/reference/src/mmr/merkle_mountain_ranges.rs:96,111,125,238-249), other mappings.
//
// Why.  A tree level with few nodes, a proof path, a peak bag are chains of DEPENDENT permutations: what counts is the
// latency of one permutation, and a B200 sub-partition issues a warp's instructions essentially one pipe at a time
// (DESIGN.md 4.2), so that latency is the length of the warp's own instruction stream.  One state per thread: 15.5 k
// instructions, 33 us for a lone warp.  Two forms split a state over threads; both run the MDS layer on the fp64 pipe
// like the thread-per-state form: the owner of an element converts its 32-bit halves to doubles (I2F) and publishes them
// in a per-warp shared-memory exchange buffer, one __syncwarp, and every thread accumulates ITS output rows over the 12
// published elements, read in ROTATED order (slot i = element (first + i) mod 12; the first elements are stored twice so
// the rotation needs no address arithmetic) -- with that order the coefficient of slot i is the same for every thread, a
// uniform operand.  Accumulators start at 2^52 + round constant, so the mantissa is the integer (combine_magic).  The
// buffer is double-buffered: one __syncwarp per round.
//
//   Quad (4 threads x 3 elements, 8 states per warp): thread j holds elements j, 4 + j, 8 + j.  Three independent S-boxes
//     per thread in a full round, 72 DFMAs per thread and layer; 870 warp instructions per state -- the THROUGHPUT form of
//     the cooperative kernels (levels of 2^5 .. 2^13 nodes, proof batches): 0.53 G permutations/s, 11 us per permutation.
//   Wide (16 lanes x 1 element, 12 used, 2 states per warp): one S-box and 24 DFMAs per lane and round; 1 900 warp
//     instructions per state but the shortest instruction stream per warp -- the LATENCY form (levels of <= 16 nodes per
//     block, the top of a tree, bagging): ~5 us per permutation.
//   (Round 1's 16-lane form used 22 warp shuffles and 24 chained IMAD.WIDE per round: 6.3 us, 2 120 instructions per state.)
//
// Partial rounds, both forms: the linear part of lanes 1..11 does not depend on the S-box of lane 0.  The owner of
// element 0 publishes ZERO for it, every thread accumulates its rows over the other eleven elements while the S-box chain
// runs (the S-box is written after the barrier so that both land in one basic block and ptxas interleaves them), then
// x = sbox(s0) is broadcast with two shuffles and enters with the per-thread coefficients M[row][0].
//
// Exactness: products of a 32-bit half with a coefficient < 64 summed over 12 lanes plus the diagonal term and a 32-bit
// constant half stay below 2^42 -- exact in the 53-bit mantissa (the same bound as the matrix-form fp64 layer of round 1);
// all values are integers, so the order of the additions does not matter.
#pragma once
#include "poseidon.cuh"

namespace poseidon {
namespace coop {

constexpr int QUAD_STRIDE = 15;                 // 16-byte slots per state: elements 0..11, then copies of 0..2 (odd: no bank conflicts)
constexpr int WIDE_STRIDE = 24;                 // elements 0..11 twice
constexpr int WARP_SLOTS = 8 * QUAD_STRIDE;     // one exchange buffer of one warp (the Wide form uses 2 * 24 of the 120)

template <int WARPS>
struct alignas(16) Shared {
  double2 xch[WARPS][2][WARP_SLOTS];       // exchange buffers (lo half, hi half of an element as doubles)
  double rc_dm[2 * WIDTH * PMT_ROUNDS];    // PMT_RC_DM: 2^52 + halves of the constants the layer of round r adds
  uint64_t rc0[WIDTH];                     // first constant layer
};

// all threads of the block; ends with a block barrier.  (Per-lane indices into constant memory serialise, but this runs
// once per block: ~1 us.)
template <int WARPS>
__device__ __forceinline__ void stage(Shared<WARPS>& sh) {
  for (int i = threadIdx.x; i < 2 * WIDTH * PMT_ROUNDS; i += blockDim.x) sh.rc_dm[i] = PMT_RC_DM[i];
  if (threadIdx.x < WIDTH) sh.rc0[threadIdx.x] = PMT_RC[threadIdx.x];
  __syncthreads();
}

__device__ __forceinline__ double2 halves(uint64_t x) { return make_double2((double)gl::lo32(x), (double)gl::hi32(x)); }

// ---------------------------------------------------------------------------------------------------------------
// Quad: thread j of 4 adjacent lanes holds e[a] = element j + 4a
// ---------------------------------------------------------------------------------------------------------------
struct Quad {
  static constexpr int LANES = 4, ELEMS = 3;
  unsigned j;            // position in the quad
  unsigned lane0;        // warp lane of the quad's thread 0
  double2* slot;         // &xch[warp][0][quad * 15 + j]: writes at +0 (+12 for j < 3), +4, +8; reads at +0 .. +11
  const double* kdm;     // rc_dm + 2 j: the constants of element j + 4a in round r are kdm[24 r + 8 a + {0, 1}]
  double c00;            // coefficient of slot 0 in output row j: C[0] + DIAG[0] = 25 on thread 0, C[0] = 17 elsewhere
  double cx[3];          // M[j + 4a][0]: coefficient of element 0 in this thread's three output rows

  template <int WARPS>
  __device__ __forceinline__ static Quad make(Shared<WARPS>& sh) {
    Quad t;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    t.j = lane & 3;
    t.lane0 = lane & ~3u;
    t.slot = &sh.xch[warp][0][(lane >> 2) * QUAD_STRIDE + t.j];
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
    for (int a = 0; a < 3; a++) e[a] = gl::add_canonical(e[a], sh.rc0[j + 4 * a]);
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
          const double xl = (double)__shfl_sync(0xffffffffu, gl::lo32(x), lane0);
          const double xh = (double)__shfl_sync(0xffffffffu, gl::hi32(x), lane0);
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
// Wide: lane g of 16 adjacent lanes holds e[0] = element g (g < 12; lanes 12..15 carry nothing and publish nothing)
// ---------------------------------------------------------------------------------------------------------------
struct Wide {
  static constexpr int LANES = 16, ELEMS = 1;
  unsigned g;            // position in the group; g >= 12: idle
  unsigned lane0;        // warp lane of the group's lane 0
  double2* slot;         // &xch[warp][0][group * 24 + g] (idle lanes: + 0): writes at +0 and +12, reads at +0 .. +11
  const double* kdm;     // rc_dm + 2 g
  double c0;             // coefficient of slot 0 (element g) in output row g: 25 on lane 0, 17 elsewhere
  double cx;             // M[g][0]

  template <int WARPS>
  __device__ __forceinline__ static Wide make(Shared<WARPS>& sh) {
    Wide t;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    t.g = lane & 15;
    t.lane0 = lane & 16u;
    const unsigned gg = t.g < WIDTH ? t.g : 0;
    t.slot = &sh.xch[warp][0][(lane >> 4) * WIDE_STRIDE + gg];
    t.kdm = sh.rc_dm + 2 * gg;
    t.c0 = t.g == 0 ? 25.0 : 17.0;
    t.cx = PMT_MDS_CIRC_D[gg == 0 ? 12 : 12 - gg];
    return t;
  }
  __device__ __forceinline__ unsigned elem(int) const { return g; }
  __device__ __forceinline__ unsigned state() const { return threadIdx.x >> 4; }
  __device__ __forceinline__ bool all(bool pred) const {
    const unsigned b = __ballot_sync(0xffffffffu, pred), mask = 0xFFFFu << lane0;
    return (b & mask) == mask;
  }

  // this lane's row over the 12 published slots: two independent chains per half (the additions are exact integers)
  __device__ __forceinline__ void row(const double2* __restrict__ rd, const double* __restrict__ k, double& L, double& H) const {
    const double2 c = *reinterpret_cast<const double2*>(k);
    double La = c.x, Ha = c.y, Lb = 0.0, Hb = 0.0;
#pragma unroll
    for (int i = 0; i < WIDTH; i += 2) {
      const double2 v = rd[i], u = rd[i + 1];
      const double cv = i == 0 ? c0 : PMT_MDS_CIRC_D[i], cu = PMT_MDS_CIRC_D[i + 1];
      La = fma(v.x, cv, La); Ha = fma(v.y, cv, Ha);
      Lb = fma(u.x, cu, Lb); Hb = fma(u.y, cu, Hb);
    }
    L = La + Lb; H = Ha + Hb;
  }

  template <int WARPS>
  __device__ __forceinline__ void permute(uint64_t (&e)[1], const Shared<WARPS>& sh) const {
    const bool on = g < WIDTH;
    uint64_t v = on ? gl::add_canonical(e[0], sh.rc0[g]) : 0ull;
    double2* s = slot;
    const double* k = kdm;
    int toggle = WARP_SLOTS;
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
#pragma unroll 1
      for (int q = 0; q < PMT_FULL_HALF; q++) {
        v = gl::pow7(v);
        const double2 d = halves(v);
        if (on) { s[0] = d; s[WIDTH] = d; }
        __syncwarp();
        double L, H;
        row(s, k, L, H);
        v = combine_magic(L, H);
        s += toggle; toggle = -toggle; k += 2 * WIDTH;
      }
      if (half == 0) {
#pragma unroll 1
        for (int p = 0; p < PMT_PARTIAL; p++) {
          const double2 d = g == 0 ? make_double2(0.0, 0.0) : halves(v);
          if (on) { s[0] = d; s[WIDTH] = d; }
          __syncwarp();
          double L, H;
          row(s, k, L, H);                              // independent of x: overlaps the S-box chain below
          const uint64_t x = gl::pow7(v);               // only lane 0's is used
          const double xl = (double)__shfl_sync(0xffffffffu, gl::lo32(x), lane0);
          const double xh = (double)__shfl_sync(0xffffffffu, gl::hi32(x), lane0);
          v = combine_magic(fma(xl, cx, L), fma(xh, cx, H));
          s += toggle; toggle = -toggle; k += 2 * WIDTH;
        }
      }
    }
    e[0] = v;
  }
};

}  // namespace coop
}  // namespace poseidon
