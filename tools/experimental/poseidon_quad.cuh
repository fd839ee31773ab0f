// EXPERIMENTAL (tools/ab_level.cu -DPMT_QUAD=1 only): the fp64 tensor-pipe form of round 1, bit-exact, not faster.
// poseidon_quad.cuh -- width-12 Poseidon over Goldilocks, 32 states per WARP, MDS layers on the fp64 tensor pipe (DMMA).
//
// Same function as poseidon.cuh's permute_paired ([UPSTREAM plonky2 hash/poseidon.rs Poseidon::poseidon], the
// arithmetic under PoseidonHash::{two_to_one, hash_or_noop} at /root/reference/src/simple_merkle_tree/
// simple_merkle_tree.rs:23,33,45 and /root/reference/src/mmr/merkle_mountain_ranges.rs:96,111,125), other mapping.
//
// Why.  Measured on B200 (tools/perm_bench.cu, tools/dmma_bench.cu; profiles/pipes_r1.jsonl, profiles/dmma_r1.jsonl):
//   * a DFMA is one multiply-accumulate per lane but holds the issue port: DFMA + LOP3 pairs take 3.7 cycles per SMSP,
//     so the 6 000 DFMAs of the thread-per-state permutation cost ~16 k of its ~30 k cycles per warp;
//   * DMMA.8x8x4 does 256 fp64 MACs in 16.1 cycles per SMSP (the same 16 MAC/clk/SMSP as DFMA) from ONE issue slot, and
//     ALU / IMAD instructions of other warps issue underneath it (DMMA + 8 LOP3: 18.2 cycles).
// The MDS layer of a batch of states is a dense 12 x 12 matrix product, so it moves to DMMA; the S-boxes (64-bit
// modular multiplications: IMAD.WIDE + carry chains) stay on the integer pipes and now overlap with it.
//
// Quad layout.  A warp owns 32 states = 4 blocks (mb) of 8.  Thread (q = lane >> 2, j = lane & 3) holds, for each mb,
// elements j, 4 + j, 8 + j of state 8 mb + q:  e[mb][t] = state[8 mb + q].lane[j + 4 t].  With the states on the M side
// of  D[state][out] += A[state][k] * B[k][out]  (m8n8k4: A one double per thread at row q, k = j), k-step t consumes
// exactly e[mb][t], and with the output columns of B ordered (0, 4, 1, 5, 2, 6, 3, 7 | 8, -, 9, -, 10, -, 11, -) thread
// (q, j) receives output lanes j, 4 + j (block 0) and 8 + j (block 1): the outputs land in the layout of the inputs and
// NO data moves between threads in a full round.  Exact integers: 32-bit halves of each element as doubles, sums below
// 2^50, accumulators start at constant + 2^52 so the mantissa is the integer (as in poseidon.cuh).
//
// Partial rounds run in pairs (poseidon.cuh permute_paired: z = A s' + col0(M) x + K): the spare column 1 of block 1
// carries y0 = row0(M) s' + c, so one DMMA pass yields z - col0(M) x and y0; the two S-boxes of lane 0 (held by the
// j = 0 threads for 4 states each) are spread over the quad with warp shuffles so every thread computes one.
#pragma once
#include "poseidon_variants.cuh"

namespace poseidonx {

constexpr int QUAD_SLOTS = 2 * PMT_FULL_HALF + PMT_PARTIAL / 2;   // 8 full rounds + 11 pairs

// per-block tables in shared memory (lane-dependent, so they cannot be constant-bank operands)
struct alignas(16) QuadTables {
  // accumulator start values of the DMMA pass of each round slot, per j:
  // [L(j), L(4+j), H(j), H(4+j), L(8+j), L(y0), H(8+j), H(y0), pad, pad], every entry = constant half + 2^52
  double acc_init[QUAD_SLOTS][4][10];   // 8 used; 80-byte rows: the 4 rows a quarter-warp reads fall in distinct banks
  uint64_t rc0[WIDTH];   // first constant layer
  // B fragments per lane (PMT_QUAD_FRAGS_SMEM): [lane][bF 3x2 | bP 3x2 | rk 3 | pad] -- reloaded at the top of every
  // round instead of living in 30 registers for the whole kernel
  double frags[16][32];   // [fragment][lane]: consecutive lanes, consecutive banks
};

// all threads of the block; ends with a block barrier
__device__ __forceinline__ void quad_stage_tables(QuadTables& T) {
  for (int idx = threadIdx.x; idx < QUAD_SLOTS * 32; idx += blockDim.x) {
    const int slot = idx >> 5, j = (idx >> 3) & 3, c = idx & 7;
    const int half = (c >> 1) & 1;
    const bool is_y0 = (c & 5) == 5;
    const int ln = c < 4 ? ((c & 1) ? 4 + j : j) : 8 + j;
    double v;
    if (slot < PMT_FULL_HALF || slot >= PMT_FULL_HALF + PMT_PARTIAL / 2) {
      const int r = slot < PMT_FULL_HALF ? slot : slot + PMT_PARTIAL / 2;   // 0..3, 26..29
      v = is_y0 ? 4503599627370496.0 : PMT_RC_DM[2 * WIDTH * r + 2 * ln + half];
    } else {
      const int p = slot - PMT_FULL_HALF, r = PMT_FULL_HALF + 2 * p;
      v = is_y0 ? PMT_RC_DM[2 * WIDTH * r + half] : PMT_PP_K_DM[2 * WIDTH * p + 2 * ln + half];
    }
    T.acc_init[slot][j][c] = v;
  }
  if (threadIdx.x < WIDTH) T.rc0[threadIdx.x] = PMT_RC[threadIdx.x];
  __syncthreads();
}

#ifndef PMT_QUAD_FRAGS_SMEM
#define PMT_QUAD_FRAGS_SMEM 0
#endif
// a shared-memory load the compiler may not hoist out of the round loops (that would put the fragments back in registers)
__device__ __forceinline__ double lds_f64_pinned(const double* p) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)));
  return v;
}

__device__ __forceinline__ double mds_entry(int row, int col) {   // M[row][col] of the full MDS matrix
  return (row == 0 && col == 0) ? PMT_MDS_CIRC_D[12] : PMT_MDS_CIRC_D[(col - row + WIDTH) % WIDTH];
}

// B fragments (one double per thread and DMMA): bF = full-round matrix M, bP = [A ; row0(M)] of the paired partial
// rounds, rk = column 0 of M at this thread's lanes (the rank-1 term of the second S-box).  k = 4 ks + j, column q.
struct QuadFrags { double bF[3][2], bP[3][2], rk[3]; };
__device__ __forceinline__ void quad_load_frags(QuadFrags& f, unsigned q, unsigned j) {
#pragma unroll
  for (int ks = 0; ks < 3; ks++) {
    const int col = 4 * ks + (int)j;
    const int row0 = (q & 1) ? 4 + (int)(q >> 1) : (int)(q >> 1);
    f.bF[ks][0] = mds_entry(row0, col);
    f.bP[ks][0] = PMT_PP_A_D[WIDTH * row0 + col];
    const int row1 = 8 + (int)(q >> 1);
    f.bF[ks][1] = (q & 1) ? 0.0 : mds_entry(row1, col);
    f.bP[ks][1] = (q & 1) ? (q == 1 ? mds_entry(0, col) : 0.0) : PMT_PP_A_D[WIDTH * row1 + col];
  }
#pragma unroll
  for (int t = 0; t < 3; t++) f.rk[t] = mds_entry((int)j + 4 * t, 0);
}

// warp 0 of the block publishes its fragments (they depend on the lane only); call between two block barriers
__device__ __forceinline__ void quad_publish_frags(QuadTables& T, const QuadFrags& f, unsigned lane) {
  if (threadIdx.x < 32) {
#pragma unroll
    for (int ks = 0; ks < 3; ks++) {
      T.frags[2 * ks][lane] = f.bF[ks][0]; T.frags[2 * ks + 1][lane] = f.bF[ks][1];
      T.frags[6 + 2 * ks][lane] = f.bP[ks][0]; T.frags[6 + 2 * ks + 1][lane] = f.bP[ks][1];
    }
#pragma unroll
    for (int t = 0; t < 3; t++) T.frags[12 + t][lane] = f.rk[t];
  }
  __syncthreads();
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ uint64_t shfl64(uint64_t v, unsigned src) {
  return glx::pack(__shfl_sync(0xffffffffu, glx::lo32(v), src), __shfl_sync(0xffffffffu, glx::hi32(v), src));
}

template <bool COMBINE_ALU>
__device__ __forceinline__ uint64_t quad_combine(double L, double H) {
  return COMBINE_ALU ? combine_magic_alu(L, H) : combine_magic_fma(L, H);
}

// one DMMA pass over the elements of block mb: acc = init + [e lo | e hi] x B.  Outputs: L0/H0 = lanes (j, 4+j),
// L1/H1 = (8+j, spare column)
__device__ __forceinline__ void quad_mma(const uint64_t (&x)[3], const double (&B)[3][2], const double* __restrict__ ci,
                                         double (&L0)[2], double (&H0)[2], double (&L1)[2], double (&H1)[2]) {
  const double2 c0 = *reinterpret_cast<const double2*>(ci), c1 = *reinterpret_cast<const double2*>(ci + 2);
  const double2 c2 = *reinterpret_cast<const double2*>(ci + 4), c3 = *reinterpret_cast<const double2*>(ci + 6);
  L0[0] = c0.x; L0[1] = c0.y; H0[0] = c1.x; H0[1] = c1.y; L1[0] = c2.x; L1[1] = c2.y; H1[0] = c3.x; H1[1] = c3.y;
#pragma unroll
  for (int ks = 0; ks < 3; ks++) {
    const double lo = (double)glx::lo32(x[ks]), hi = (double)glx::hi32(x[ks]);   // I2F.F64.U32 (conversion pipe)
    dmma884(L0[0], L0[1], lo, B[ks][0]);
    dmma884(H0[0], H0[1], hi, B[ks][0]);
    dmma884(L1[0], L1[1], lo, B[ks][1]);
    dmma884(H1[0], H1[1], hi, B[ks][1]);
  }
}

// The permutation of the warp's 32 states, in place.  Warp-collective: all 32 lanes must call it together.
// Output elements are NOT canonicalised.  SBOX_FMA_MASK / PART_FMA_MASK: pow7_mix masks of the full / partial rounds.
template <int SBOX_FMA_MASK = 0, int PART_FMA_MASK = 0, bool COMBINE_ALU = true>
__device__ __forceinline__ void permute_quad(uint64_t (&e)[4][3], const QuadTables& T, const QuadFrags& f, unsigned lane) {
  const unsigned j = lane & 3, quad0 = lane & ~3u;
#pragma unroll
  for (int mb = 0; mb < 4; mb++)
#pragma unroll
    for (int t = 0; t < 3; t++) e[mb][t] = glx::add_canonical(e[mb][t], T.rc0[j + 4 * t]);
#pragma unroll 1
  for (int half = 0; half < 2; half++) {
#pragma unroll 1
    for (int r = 0; r < PMT_FULL_HALF; r++) {
      const double* ci = T.acc_init[half ? PMT_FULL_HALF + PMT_PARTIAL / 2 + r : r][j];
      double bF[3][2];
#pragma unroll
      for (int ks = 0; ks < 3; ks++)
#pragma unroll
        for (int nb = 0; nb < 2; nb++) bF[ks][nb] = PMT_QUAD_FRAGS_SMEM ? lds_f64_pinned(&T.frags[2 * ks + nb][lane]) : f.bF[ks][nb];
#pragma unroll
      for (int mb = 0; mb < 4; mb++) {
        uint64_t x[3];
#pragma unroll
        for (int t = 0; t < 3; t++) x[t] = pow7_mix<SBOX_FMA_MASK>(e[mb][t]);
        double L0[2], H0[2], L1[2], H1[2];
        quad_mma(x, bF, ci, L0, H0, L1, H1);
        e[mb][0] = quad_combine<COMBINE_ALU>(L0[0], H0[0]);
        e[mb][1] = quad_combine<COMBINE_ALU>(L0[1], H0[1]);
        e[mb][2] = quad_combine<COMBINE_ALU>(L1[0], H1[0]);
      }
    }
    if (half == 0) {
#pragma unroll 1
      for (int pair = 0; pair < PMT_PARTIAL / 2; pair++) {
        const double* ci = T.acc_init[PMT_FULL_HALF + pair][j];
        double bP[3][2], rk[3];
#pragma unroll
        for (int ks = 0; ks < 3; ks++)
#pragma unroll
          for (int nb = 0; nb < 2; nb++) bP[ks][nb] = PMT_QUAD_FRAGS_SMEM ? lds_f64_pinned(&T.frags[6 + 2 * ks + nb][lane]) : f.bP[ks][nb];
        // first S-box: lane 0 of state 8 m + q sits in thread (q, 0); thread (q, m) computes it
        uint64_t v = e[0][0];
#pragma unroll
        for (int m = 1; m < 4; m++) { const uint64_t t = shfl64(e[m][0], quad0); v = j == (unsigned)m ? t : v; }
        v = pow7_mix<PART_FMA_MASK>(v);
#pragma unroll
        for (int m = 0; m < 4; m++) { const uint64_t t = m ? shfl64(v, quad0 | m) : v; e[m][0] = j == 0 ? t : e[m][0]; }
        // one DMMA pass: z - col0(M) x (lanes j, 4+j, 8+j) and y0 (spare column; valid in the j = 0 threads)
        double L0[4][2], H0[4][2], L1[4], H1[4];
        uint64_t y0[4];
#pragma unroll
        for (int mb = 0; mb < 4; mb++) {
          double l1[2], h1[2];
          quad_mma(e[mb], bP, ci, L0[mb], H0[mb], l1, h1);
          L1[mb] = l1[0]; H1[mb] = h1[0];
          y0[mb] = quad_combine<COMBINE_ALU>(l1[1], h1[1]);
        }
        // second S-box, spread over the quad the same way; then every thread needs x of its 4 states
        v = y0[0];
#pragma unroll
        for (int m = 1; m < 4; m++) { const uint64_t t = shfl64(y0[m], quad0); v = j == (unsigned)m ? t : v; }
        v = pow7_mix<PART_FMA_MASK>(v);
#pragma unroll
        for (int t = 0; t < 3; t++) rk[t] = PMT_QUAD_FRAGS_SMEM ? lds_f64_pinned(&T.frags[12 + t][lane]) : f.rk[t];
#pragma unroll
        for (int m = 0; m < 4; m++) {
          const uint64_t x = shfl64(v, quad0 | m);
          const double xlo = (double)glx::lo32(x), xhi = (double)glx::hi32(x);
          e[m][0] = quad_combine<COMBINE_ALU>(fma(xlo, rk[0], L0[m][0]), fma(xhi, rk[0], H0[m][0]));
          e[m][1] = quad_combine<COMBINE_ALU>(fma(xlo, rk[1], L0[m][1]), fma(xhi, rk[1], H0[m][1]));
          e[m][2] = quad_combine<COMBINE_ALU>(fma(xlo, rk[2], L1[m]), fma(xhi, rk[2], H1[m]));
        }
      }
    }
  }
}

}  // namespace poseidonx
