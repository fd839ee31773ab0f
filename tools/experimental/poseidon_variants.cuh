// poseidon_variants.cuh -- EXPERIMENTAL (tools/ab_level.cu only; the product uses csrc/poseidon.cuh): the permutation forms
// built and measured in round 1 (specification form, sparse partial rounds, IMAD.WIDE / limb / DFMA MDS layers, fused,
// paired, paired + frequency domain with all its knobs, 16-lane cooperative), in namespace poseidonx.
// poseidon.cuh -- width-12 Poseidon permutation over Goldilocks, one state per thread, in registers.
//
// Replaces [UPSTREAM plonky2 hash/poseidon.rs Poseidon::poseidon, hash/poseidon_goldilocks.rs, hash/hashing.rs
// compress / hash_n_to_m_no_pad], i.e. what PoseidonHash::{two_to_one, hash_or_noop} execute for
// /root/reference/src/simple_merkle_tree/simple_merkle_tree.rs:23,33,45 and
// /root/reference/src/mmr/merkle_mountain_ranges.rs:96,111,125.
//
// sm_100a mapping (no tensor cores: this is not a contraction):
//   * state = 12 x u64 = 24 registers per thread; rounds fully unrolled so every round constant is a c[3][imm] operand.
//   * MDS layer: the circulant coefficients are < 64, so each lane is split in 32-bit halves and every
//     coefficient*half MAC is ONE IMAD.WIDE.U32 into a 64-bit column accumulator (no carries: sums < 2^42);
//     the next round's constants are the accumulators' initial values, so "add round constants" costs nothing.
//     Two column sums are folded with 2^64 = 2^32 - 1 (glx::combine_halves, 1 IMAD + 4 ALU).
//   * S-box x^7 = 2 squarings + 2 multiplies, 16 SASS instructions each (glx::mul).
#pragma once
#include "goldilocks_variants.cuh"
#include "../../plonky2_merkle_trees_b200/csrc/poseidon_constants.cuh"
#include "../../plonky2_merkle_trees_b200/csrc/poseidon_freq.cuh"

#ifndef PMT_COMBINE_C
#define PMT_COMBINE_C 1   // 1: the ALU recombination is combine_magic_c (plain C, 3-input adds) instead of the PTX carry chain
#endif

namespace poseidonx {
namespace freq = poseidon::freq;

static constexpr int WIDTH = 12;

// out[r] = add[r] + sum_i s[(i + r) % 12] * CIRC[i] + 8 * s[0] * [r == 0];   add = next round's constants
//
// Measured on B200 (tools/perm_bench.cu, profiles/pipes_r1.jsonl): IMAD.WIDE.U32 with a zero addend issues at
// 64 lanes/clk/SM, but with a 64-bit register addend (the "accumulate" form) only at 32, and IMAD.HI at 32.  So the
// MACs are written as plain C: ptxas turns them into zero-addend IMAD.WIDE.U32 (fma pipe) plus 3-input
// IADD3/IADD3.X pairs that fold two products per pair (alu pipe) -- one slot on each pipe per product, which
// balances against the S-box (also one fma slot per alu slot).
__device__ __forceinline__ void mds_layer(uint64_t (&s)[WIDTH], const uint64_t* __restrict__ add) {
  uint32_t lo[WIDTH], hi[WIDTH], coef[WIDTH];
#pragma unroll
  for (int i = 0; i < WIDTH; i++) { lo[i] = glx::lo32(s[i]); hi[i] = glx::hi32(s[i]); coef[i] = PMT_MDS_CIRC32[i]; }
  const uint32_t coef00 = PMT_MDS_CIRC32[12];
#pragma unroll
  for (int r = 0; r < WIDTH; r++) {
    // low column starts at the full 64-bit round constant: every constant is < 2^64 - 2^48 (upstream invariant,
    // asserted by tools/gen_constants.py) and the column sums are < 2^42, so nothing overflows
    uint64_t L = add ? add[r] : 0ull;
    uint64_t H = 0ull;
#pragma unroll
    for (int i = 0; i < WIDTH; i++) {
      const uint32_t c = (r == 0 && i == 0) ? coef00 : coef[i];  // MDS_MATRIX_DIAG = [8, 0, ..., 0]
      L += (uint64_t)lo[(i + r) % WIDTH] * c;
      H += (uint64_t)hi[(i + r) % WIDTH] * c;
    }
    s[r] = glx::combine_halves(L, H);
  }
}

// Limb form of the same layer.  Each lane is split 22/21/21 bits; limb * coefficient sums (12 terms + the constant's
// limb) stay below 2^31, so every MAC is one full-rate 32-bit IMAD with a free accumulate (432 fma slots instead of the
// 576 that 288 half-rate IMAD.WIDE accumulates cost), at the price of 4 ALU ops per lane to split and 15 to recombine
// y = A0 + 2^22 A1 + 2^43 A2 (74 bits) and fold the top word with 2^64 = 2^32 - 1.  add_l = limbs of the constants.
__device__ __forceinline__ void mds_layer_limb3(uint64_t (&s)[WIDTH], const uint32_t* __restrict__ add_l) {
  uint32_t l0[WIDTH], l1[WIDTH], l2[WIDTH], coef[WIDTH];
#pragma unroll
  for (int i = 0; i < WIDTH; i++) {
    const uint32_t w0 = glx::lo32(s[i]), w1 = glx::hi32(s[i]);
    l0[i] = w0 & 0x3FFFFFu;
    l1[i] = __funnelshift_r(w0, w1, 22) & 0x1FFFFFu;
    l2[i] = w1 >> 11;
    coef[i] = PMT_MDS_CIRC32[i];
  }
  const uint32_t coef00 = PMT_MDS_CIRC32[12];
#pragma unroll
  for (int r = 0; r < WIDTH; r++) {
    uint32_t a0 = add_l[3 * r], a1 = add_l[3 * r + 1], a2 = add_l[3 * r + 2];
#pragma unroll
    for (int i = 0; i < WIDTH; i++) {
      const uint32_t c = (r == 0 && i == 0) ? coef00 : coef[i];
      a0 += l0[(i + r) % WIDTH] * c;
      a1 += l1[(i + r) % WIDTH] * c;
      a2 += l2[(i + r) % WIDTH] * c;
    }
    uint32_t lo, hi, top, net;
    // (top:hi:lo) = a0 + (a1 << 22) + (a2 << 43)
    asm("add.cc.u32 %0, %3, %4;\n\taddc.cc.u32 %1, %5, %6;\n\taddc.u32 %2, %7, 0;"
        : "=r"(lo), "=r"(hi), "=r"(top) : "r"(a0), "r"(a1 << 22), "r"(a1 >> 10), "r"(a2 << 11), "r"(a2 >> 21));
    // + top * (2^32 - 1) = (top << 32) - top; net = carry - borrow is 0 or 1 (a borrow forces the carry, as in
    // glx::reduce128_alu), folded once more
    asm("sub.cc.u32 %0, %0, %3;\n\tsubc.cc.u32 %1, %1, 0;\n\tsubc.u32 %2, 0, 0;" : "+r"(lo), "+r"(hi), "=r"(net) : "r"(top));
    asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, 0;" : "+r"(hi), "+r"(net) : "r"(top));
    asm("sub.cc.u32 %0, %0, %2;\n\tsubc.u32 %1, %1, 0;" : "+r"(lo), "+r"(hi) : "r"(net));
    hi += net;
    s[r] = glx::pack(lo, hi);
  }
}

// fp64 form of the same layer.  A DFMA is a multiply-accumulate in ONE fma-pipe slot (the accumulate form of
// IMAD.WIDE.U32 takes two), and integers below 2^52 are exact in a double: the 32-bit halves are converted with the
// 2^52 magic-number trick (pair the word with 0x43300000, subtract 2^52), the 12-term column sums stay below 2^42, and
// adding 2^52 back leaves the integer in the mantissa.  add_d = constants as (lo, hi) doubles.
// rows_4_only (warp-uniform): compute output lanes 0..3 only -- the last layer of a permutation whose caller keeps just
// the digest (two_to_one, last sponge block); lanes 4..11 are left undefined.
template <bool CVT_I2F = false>
__device__ __forceinline__ void mds_layer_dfma(uint64_t (&s)[WIDTH], const double* __restrict__ add_d,
                                               bool rows_4_only = false) {
  const double MAGIC = 4503599627370496.0;  // 2^52
  double dlo[WIDTH], dhi[WIDTH], coef[WIDTH];
#pragma unroll
  for (int i = 0; i < WIDTH; i++) {
    if (CVT_I2F) {   // I2F.F64.U32 runs on the (otherwise idle) conversion pipe instead of MOV + DADD on the fma pipe
      dlo[i] = (double)glx::lo32(s[i]);
      dhi[i] = (double)glx::hi32(s[i]);
    } else {
      dlo[i] = __hiloint2double(0x43300000, (int)glx::lo32(s[i])) - MAGIC;
      dhi[i] = __hiloint2double(0x43300000, (int)glx::hi32(s[i])) - MAGIC;
    }
    coef[i] = PMT_MDS_CIRC_D[i];
  }
  const double coef00 = PMT_MDS_CIRC_D[12];
  uint64_t out[WIDTH];
#pragma unroll
  for (int r = 0; r < WIDTH; r++) {
    if (r == 4 && rows_4_only) break;
    double L = add_d[2 * r], H = add_d[2 * r + 1];
#pragma unroll
    for (int i = 0; i < WIDTH; i++) {
      const double c = (r == 0 && i == 0) ? coef00 : coef[i];
      L = fma(dlo[(i + r) % WIDTH], c, L);
      H = fma(dhi[(i + r) % WIDTH], c, H);
    }
    const uint64_t li = (uint64_t)__double_as_longlong(L + MAGIC) & 0x000FFFFFFFFFFFFFull;
    const uint64_t hi = (uint64_t)__double_as_longlong(H + MAGIC) & 0x000FFFFFFFFFFFFFull;
    out[r] = glx::combine_halves(li, hi);
  }
#pragma unroll
  for (int r = 0; r < WIDTH; r++) s[r] = out[r];
}

// Specification form: 30 x (add constants, x^7 on all lanes / lane 0, MDS).  Output lanes are NOT canonicalised.
// One rolled loop over the rounds: the body (12 s-boxes + 1 s-box + one MDS layer, ~1.2k SASS instructions = 19 KB)
// stays inside the 32 KB L1.5 instruction cache; a fully unrolled permutation (~290 KB) would be fetch-bound.
template <bool SBOX_ALU = false>
__device__ __forceinline__ void permute_naive(uint64_t (&s)[WIDTH]) {
#pragma unroll
  for (int i = 0; i < WIDTH; i++) s[i] = glx::add_canonical(s[i], PMT_RC[i]);
#pragma unroll 1
  for (int r = 0; r < PMT_ROUNDS; r++) {
    if (r < PMT_FULL_HALF || r >= PMT_FULL_HALF + PMT_PARTIAL) {
#pragma unroll
      for (int i = 0; i < WIDTH; i++) s[i] = glx::pow7<SBOX_ALU>(s[i]);
    } else {
      s[0] = glx::pow7<SBOX_ALU>(s[0]);
    }
    // the table has 30 rows; the last layer adds row 30 = zeros (PMT_RC_PAD)
    mds_layer(s, &PMT_RC[WIDTH * (r + 1)]);
  }
}

// sum_{t < N} x[t] * K[t] mod p, K given as 22/22/20-bit limbs kl[3 t + j] (constant memory).
// Each 32-bit half of x[t] times a limb is < 2^54; N <= 12 such terms stay < 2^58, so the six column sums need no
// carry handling at all (plain 64-bit adds, which ptxas pairs into 3-input IADD3/IADD3.X).
template <int N>
__device__ __forceinline__ uint64_t dot_limbs(const uint64_t (&x)[N], const uint32_t* __restrict__ kl) {
  uint64_t t00 = 0, t01 = 0, t02 = 0, t10 = 0, t11 = 0, t12 = 0;
#pragma unroll
  for (int t = 0; t < N; t++) {
    const uint32_t a0 = glx::lo32(x[t]), a1 = glx::hi32(x[t]);
    const uint32_t b0 = kl[3 * t], b1 = kl[3 * t + 1], b2 = kl[3 * t + 2];
    // glx::mad_wide (the mad.lo.cc / madc.hi spelling) keeps these as bare chained IMAD.WIDE.U32; the plain C form made
    // ptxas add one junk VIADD (adding a zero uniform register) per product, all on the bottleneck fma pipe
    t00 = glx::mad_wide(a0, b0, t00); t01 = glx::mad_wide(a0, b1, t01); t02 = glx::mad_wide(a0, b2, t02);
    t10 = glx::mad_wide(a1, b0, t10); t11 = glx::mad_wide(a1, b1, t11); t12 = glx::mad_wide(a1, b2, t12);
  }
  // V = G0 + 2^32 G1,  G0 = t00 + 2^22 t01 + 2^44 t02 (< 2^103), G1 likewise.
  // 2^32 G1 = 2^32 g_lo + 2^96 g_hi = 2^32 g_lo - g_hi (mod p); adding 2^40 p keeps the total non-negative.
  glx::u128 g0 = (glx::u128)t00 + ((glx::u128)t01 << 22) + ((glx::u128)t02 << 44);
  glx::u128 g1 = (glx::u128)t10 + ((glx::u128)t11 << 22) + ((glx::u128)t12 << 44);
  const uint64_t g_lo = (uint64_t)g1, g_hi = (uint64_t)(g1 >> 64);
  glx::u128 v = g0 + ((glx::u128)g_lo << 32) + (((glx::u128)glx::P << 40) - g_hi);
  return glx::reduce128(v);
}

// The same permutation with the 22 partial rounds in the sparse-matrix ("fast") form that upstream also executes
// [UPSTREAM hash/poseidon.rs partial_first_constant_layer / mds_partial_layer_init / mds_partial_layer_fast]; tables
// re-derived in tools/gen_constants.py.  Per partial round: 1 s-box, one 12-term dot product, 11 multiply-adds.
//
// Code layout matters as much as instruction count here: ncu on the first version (two copies of the full round,
// 56 KB of SASS) showed a 93.7 % instruction-cache hit rate and 0.45 "no instruction" stalls per issue.  The two
// halves therefore share ONE full-round body (outer loop over the halves) and the dot product is one shared
// __noinline__ routine, which keeps the whole permutation under the 32 KB L1.5 instruction cache.
// (scalar parameters: the CUDA ABI passes them in registers, an array reference would spill the state to local memory)
__device__ __noinline__ uint64_t dot12_limbs(uint64_t x0, uint64_t x1, uint64_t x2, uint64_t x3, uint64_t x4, uint64_t x5,
                                             uint64_t x6, uint64_t x7, uint64_t x8, uint64_t x9, uint64_t x10,
                                             uint64_t x11, const uint32_t* __restrict__ kl) {
  const uint64_t x[WIDTH] = {x0, x1, x2, x3, x4, x5, x6, x7, x8, x9, x10, x11};
  return dot_limbs<WIDTH>(x, kl);
}
__device__ __forceinline__ uint64_t dot12_limbs(const uint64_t (&s)[WIDTH], const uint32_t* __restrict__ kl) {
  return dot12_limbs(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8], s[9], s[10], s[11], kl);
}

// x^7 on four lanes as ONE shared routine (3 calls per full round): the unrolled 12-lane S-box layer is 12 KB of SASS,
// this is 4 KB, which keeps the hot loops inside the instruction cache.  SBOX_CALL selects it.
template <bool ALU>
__device__ __noinline__ void pow7_x4(uint64_t& a, uint64_t& b, uint64_t& c, uint64_t& d) {
  a = glx::pow7<ALU>(a); b = glx::pow7<ALU>(b); c = glx::pow7<ALU>(c); d = glx::pow7<ALU>(d);
}

// MDS_MODE: 0 = 64-bit column sums (IMAD.WIDE), 1 = 22/21/21-bit limbs (32-bit IMAD), 2 = fp64 column sums (DFMA),
//           3 = fp64 column sums with I2F conversions
// CAP_ZERO: the caller guarantees lanes 8..11 are zero on entry (two_to_one): their first S-box is a table lookup.
// OUT4: the caller only reads lanes 0..3 afterwards: the last MDS layer computes 4 rows (MDS_MODE 2 only).
template <bool SBOX_ALU = false, bool PART_ALU = false, int MDS_MODE = 0, bool CAP_ZERO = false, bool OUT4 = false,
          bool SBOX_CALL = false>
__device__ __forceinline__ void permute_fast(uint64_t (&s)[WIDTH]) {
#pragma unroll
  for (int i = 0; i < (CAP_ZERO ? 8 : WIDTH); i++) s[i] = glx::add_canonical(s[i], PMT_RC[i]);
#pragma unroll 1
  for (int half = 0; half < 2; half++) {
#pragma unroll 1
    for (int r = 0; r < PMT_FULL_HALF; r++) {
      if (SBOX_CALL) {
        pow7_x4<SBOX_ALU>(s[0], s[1], s[2], s[3]);
        pow7_x4<SBOX_ALU>(s[4], s[5], s[6], s[7]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = glx::pow7<SBOX_ALU>(s[i]);
      }
      if (SBOX_CALL && !CAP_ZERO) {
        pow7_x4<SBOX_ALU>(s[8], s[9], s[10], s[11]);
      } else if (CAP_ZERO && half == 0 && r == 0) {
#pragma unroll
        for (int i = 8; i < WIDTH; i++) s[i] = PMT_SBOX_RC_CAP[i - 8];
      } else {
#pragma unroll
        for (int i = 8; i < WIDTH; i++) s[i] = glx::pow7<SBOX_ALU>(s[i]);
      }
      if (MDS_MODE == 1) mds_layer_limb3(s, &PMT_RC_AFTER_FULL_L[3 * WIDTH * (PMT_FULL_HALF * half + r)]);
      else if (MDS_MODE == 2) mds_layer_dfma<false>(s, &PMT_RC_AFTER_FULL_D[2 * WIDTH * (PMT_FULL_HALF * half + r)],
                                                    OUT4 && half == 1 && r == PMT_FULL_HALF - 1);
      else if (MDS_MODE == 3) mds_layer_dfma<true>(s, &PMT_RC_AFTER_FULL_D[2 * WIDTH * (PMT_FULL_HALF * half + r)],
                                                   OUT4 && half == 1 && r == PMT_FULL_HALF - 1);
      else mds_layer(s, &PMT_RC_AFTER_FULL[WIDTH * (PMT_FULL_HALF * half + r)]);
    }
    if (half == 0) {
      {  // dense INIT matrix on lanes 1..11 (lane 0 passes through); rows carry a zero coefficient for lane 0
        uint64_t y[WIDTH];
#pragma unroll 1
        for (int a = 1; a < WIDTH; a++) {
          const uint64_t v = dot12_limbs(s, &PMT_FP_INIT_L[3 * WIDTH * (a - 1)]);
#pragma unroll
          for (int i = 1; i < WIDTH; i++) if (i == a) y[i] = v;   // static indexing keeps y[] in registers
        }
#pragma unroll
        for (int i = 1; i < WIDTH; i++) s[i] = y[i];
      }
#pragma unroll 1
      for (int k = 0; k < PMT_PARTIAL; k++) {
        s[0] = glx::add_canonical(glx::pow7<PART_ALU>(s[0]), PMT_FP_POST_RC[k]);
        const uint64_t d = dot12_limbs(s, &PMT_FP_W_HAT_L[3 * WIDTH * k]);
#pragma unroll
        for (int i = 1; i < WIDTH; i++) s[i] = glx::mul_add<PART_ALU>(s[0], PMT_FP_V[(WIDTH - 1) * k + (i - 1)], s[i]);
        s[0] = d;
      }
#pragma unroll
      for (int i = 0; i < WIDTH; i++) s[i] = glx::add_canonical(s[i], PMT_RC[WIDTH * (PMT_FULL_HALF + PMT_PARTIAL) + i]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Fused form (production).  ncu on the phase-structured permute_fast showed the two issue pipes taking turns instead
// of overlapping: the S-box layer is ALU-heavy (carry chains of the reductions), the DFMA MDS layer is pure fma pipe,
// and with 4 warps per scheduler the phases of different warps rarely complement each other (fma-heavy pipe 61-70 %
// busy, 0.55 IPC).  Here every phase carries its own mix:
//   * full round = 12 x { S-box of lane i ; lane i's COLUMN of the MDS matrix: 24 DFMAs into the 12 (lo, hi) row
//     accumulators }.  The DFMAs of lane i have no consumer until the end of the round, so they fill the fma pipe
//     while the next lanes' reductions run on the ALU pipe.  The accumulators start at constant + 2^52, so the
//     integer result is the mantissa (no conversion back), and are recombined on the ALU pipe.
//   * partial round: everything inline in one basic block (no call ABI moves); the 11-term dot product of the old
//     lanes does not depend on the S-box of lane 0, so ptxas overlaps the S-box's serial chain with it; lane 0 enters
//     the dot product with its small coefficient M00 = 25 (2 MACs instead of 6).
// ---------------------------------------------------------------------------------------------------------------
// Scheduling fence.  ptxas orders a basic block by critical path: left alone it runs every S-box of a round first and
// all DFMAs afterwards (and, in the partial rounds, the S-box chain before the dot product), which recreates the
// phases.  tie(x, after) makes x formally depend on `after` through one LOP3 with a run-time zero (x | (after & 0)), so
// "lane i+2's S-box may not start before lane i's DFMAs" is a data dependency the scheduler has to respect, and the
// DFMAs of lane i are left to overlap with the S-box of lane i+1.
static __device__ __constant__ uint32_t PMT_ZERO32 = 0;   // not const: the compiler must treat it as unknown
__device__ __forceinline__ uint64_t tie(uint64_t x, uint32_t after, uint32_t zero) {
  uint32_t lo = glx::lo32(x);
  asm("lop3.b32 %0, %0, %1, %2, 0xf8;" : "+r"(lo) : "r"(after), "r"(zero));   // lo | (after & zero)
  return glx::pack(lo, glx::hi32(x));
}
__device__ __forceinline__ uint64_t tie_hi(uint64_t x, uint32_t after, uint32_t zero) {   // the same fence on the high word
  uint32_t hi = glx::hi32(x);
  asm("lop3.b32 %0, %0, %1, %2, 0xf8;" : "+r"(hi) : "r"(after), "r"(zero));
  return glx::pack(glx::lo32(x), hi);
}
__device__ __forceinline__ uint32_t hi_word(double d) { return (uint32_t)__double2hiint(d); }

// 2^52 + l, 2^52 + h (l, h < 2^43) -> u64 congruent to l + 2^32 h, ALU pipe only.
__device__ __forceinline__ uint64_t combine_magic_c(double L, double H);
__device__ __forceinline__ uint64_t combine_magic_alu(double L, double H) {
#if PMT_COMBINE_C
  return combine_magic_c(L, H);
#endif
  const uint64_t lb = (uint64_t)__double_as_longlong(L), hb = (uint64_t)__double_as_longlong(H);
  uint32_t w0 = glx::lo32(lb), w1 = glx::lo32(hb), w2 = glx::hi32(hb) & 0xFFFFFu, net;
  const uint32_t lh = glx::hi32(lb) & 0xFFFFFu;
  // (w2 : w1 : w0) = l + 2^32 h
  asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, 0;" : "+r"(w1), "+r"(w2) : "r"(lh));
  // + w2 * (2^32 - 1) = (w2 << 32) - w2; net = carry - borrow is 0 or 1 (a borrow forces the carry, see glx::reduce128_alu)
  asm("sub.cc.u32 %0, %0, %3;\n\tsubc.cc.u32 %1, %1, 0;\n\tsubc.u32 %2, 0, 0;" : "+r"(w0), "+r"(w1), "=r"(net) : "r"(w2));
  asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, 0;" : "+r"(w1), "+r"(net) : "r"(w2));
  asm("sub.cc.u32 %0, %0, %2;\n\tsubc.u32 %1, %1, 0;" : "+r"(w0), "+r"(w1) : "r"(net));
  w1 += net;
  return glx::pack(w0, w1);
}
// The same recombination in plain C (signed two-word form, like glx::reduce128_c): V = (l0 - h1) + 2^32 (l1 + h0 + h1);
// the exponent bits 0x43300000 of both high words are removed by constants folded into the 3-input adds, so there is
// no masking.  Exact for every l, h < 2^52 (tools/check_combine.c).
__device__ __forceinline__ uint64_t combine_magic_c(double L, double H) {
  const uint64_t lb = (uint64_t)__double_as_longlong(L), hb = (uint64_t)__double_as_longlong(H);
  const int64_t lo = (int64_t)(uint64_t)glx::lo32(lb) - (int64_t)(uint64_t)glx::hi32(hb) + 0x43300000ll;
  const int64_t hi = (int64_t)(uint64_t)glx::hi32(lb) + (int64_t)(uint64_t)glx::lo32(hb) + (int64_t)(uint64_t)glx::hi32(hb) +
                     (lo >> 32) - 0x86600000ll;
  const int64_t n = hi >> 32;
  return glx::pack((uint32_t)lo, (uint32_t)hi) + ((uint64_t)n << 32) - (uint64_t)n;
}
__device__ __forceinline__ uint64_t combine_magic_fma(double L, double H) {
  const uint64_t li = (uint64_t)__double_as_longlong(L) & 0x000FFFFFFFFFFFFFull;
  const uint64_t hi = (uint64_t)__double_as_longlong(H) & 0x000FFFFFFFFFFFFFull;
  return glx::combine_halves(li, hi);
}

// x^7 with a per-multiplication choice of reduction: bit j of FMA_MASK = multiplication j (x2, x4, x3, x7) folds with
// IMAD.WIDE (fma pipe, 6 ALU + 2 IMAD.WIDE) instead of the 13-instruction ALU-only reduction.
template <int FMA_MASK>
__device__ __forceinline__ uint64_t pow7_mix(uint64_t x) {
  const uint64_t x2 = (FMA_MASK & 1) ? glx::sqr<false>(x) : glx::sqr<true>(x);
  const uint64_t x4 = (FMA_MASK & 2) ? glx::sqr<false>(x2) : glx::sqr<true>(x2);
  const uint64_t x3 = (FMA_MASK & 4) ? glx::mul<false>(x, x2) : glx::mul<true>(x, x2);
  return (FMA_MASK & 8) ? glx::mul<false>(x3, x4) : glx::mul<true>(x3, x4);
}

// one full round: s[] holds the state WITH this round's constants added; add_dm = the next constants as
// (lo + 2^52, hi + 2^52) doubles.  cap_const / out4 are warp-uniform run-time flags (no code duplication):
// cap_const: lanes 8..11 hold RC[8..11] (first round of two_to_one), their S-box output is a table constant;
// out4: only output lanes 0..3 are needed (last round when the caller keeps the digest).
template <int SBOX_FMA_MASK, bool CVT_I2F, bool COMBINE_ALU, int PIPE>
__device__ __forceinline__ void full_round_fused(uint64_t (&s)[WIDTH], const double* __restrict__ add_dm, bool cap_const,
                                                 bool out4) {
  const double MAGIC = 4503599627370496.0;  // 2^52
  const uint32_t zero = PMT_ZERO32;
  double L[WIDTH], H[WIDTH];
#pragma unroll
  for (int r = 0; r < WIDTH; r++) { L[r] = add_dm[2 * r]; H[r] = add_dm[2 * r + 1]; }
#pragma unroll
  for (int i = 0; i < WIDTH; i++) {
    uint64_t x;
    if (PIPE > 0 && i >= PIPE + 1) s[i] = tie(s[i], hi_word(H[WIDTH - 1]), zero);   // after lane (i - PIPE - 1)'s DFMAs
    if (i >= 8 && cap_const) x = PMT_SBOX_RC_CAP[i - 8];
    else x = pow7_mix<SBOX_FMA_MASK>(s[i]);
    double dlo, dhi;
    if (CVT_I2F) {
      dlo = (double)glx::lo32(x); dhi = (double)glx::hi32(x);
    } else {
      dlo = __hiloint2double(0x43300000, (int)glx::lo32(x)) - MAGIC;
      dhi = __hiloint2double(0x43300000, (int)glx::hi32(x)) - MAGIC;
    }
    if (out4) {
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const double c = (i == 0 && r == 0) ? PMT_MDS_CIRC_D[12] : PMT_MDS_CIRC_D[(i - r + WIDTH) % WIDTH];
        L[r] = fma(dlo, c, L[r]); H[r] = fma(dhi, c, H[r]);
      }
    } else {
#pragma unroll
      for (int r = 0; r < WIDTH; r++) {
        const double c = (i == 0 && r == 0) ? PMT_MDS_CIRC_D[12] : PMT_MDS_CIRC_D[(i - r + WIDTH) % WIDTH];
        L[r] = fma(dlo, c, L[r]); H[r] = fma(dhi, c, H[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < WIDTH; r++) s[r] = COMBINE_ALU ? combine_magic_alu(L[r], H[r]) : combine_magic_fma(L[r], H[r]);
}

// The DFMA MDS layer with the 2^52 folded into the constants (add_dm) and the ALU-only recombination.
// COLUMN = false: row by row (24 input doubles live, 2 accumulators at a time); true: lane by lane (24 accumulators).
template <bool COLUMN, bool CVT_I2F, bool COMBINE_ALU>
__device__ __forceinline__ void mds_layer_dfma2(uint64_t (&s)[WIDTH], const double* __restrict__ add_dm, bool out4) {
  const double MAGIC = 4503599627370496.0;  // 2^52
  double dlo[WIDTH], dhi[WIDTH];
#pragma unroll
  for (int i = 0; i < WIDTH; i++) {
    if (CVT_I2F) { dlo[i] = (double)glx::lo32(s[i]); dhi[i] = (double)glx::hi32(s[i]); }
    else {
      dlo[i] = __hiloint2double(0x43300000, (int)glx::lo32(s[i])) - MAGIC;
      dhi[i] = __hiloint2double(0x43300000, (int)glx::hi32(s[i])) - MAGIC;
    }
  }
  if (!COLUMN) {
    uint64_t out[WIDTH];
#pragma unroll
    for (int r = 0; r < WIDTH; r++) {
      if (r == 4 && out4) break;
      double L = add_dm[2 * r], H = add_dm[2 * r + 1];
#pragma unroll
      for (int i = 0; i < WIDTH; i++) {
        const double c = (r == 0 && i == 0) ? PMT_MDS_CIRC_D[12] : PMT_MDS_CIRC_D[i];
        L = fma(dlo[(i + r) % WIDTH], c, L);
        H = fma(dhi[(i + r) % WIDTH], c, H);
      }
      out[r] = COMBINE_ALU ? combine_magic_alu(L, H) : combine_magic_fma(L, H);
    }
#pragma unroll
    for (int r = 0; r < WIDTH; r++) s[r] = out[r];
  } else {
    double L[WIDTH], H[WIDTH];
#pragma unroll
    for (int r = 0; r < WIDTH; r++) { L[r] = add_dm[2 * r]; H[r] = add_dm[2 * r + 1]; }
    // lane 0 (the only lane behind an S-box in a partial round) goes last
#pragma unroll
    for (int ii = 1; ii <= WIDTH; ii++) {
      const int i = ii % WIDTH;
#pragma unroll
      for (int r = 0; r < WIDTH; r++) {
        if (r >= 4 && out4) break;
        const double c = (i == 0 && r == 0) ? PMT_MDS_CIRC_D[12] : PMT_MDS_CIRC_D[(i - r + WIDTH) % WIDTH];
        L[r] = fma(dlo[i], c, L[r]); H[r] = fma(dhi[i], c, H[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < WIDTH; r++) s[r] = COMBINE_ALU ? combine_magic_alu(L[r], H[r]) : combine_magic_fma(L[r], H[r]);
  }
}

// Specification form, ONE rolled loop over the 30 rounds (constants, S-box on 12 lanes / lane 0, MDS), with the DFMA
// MDS layer.  With the MDS layer at 312 fma-pipe slots a partial round costs about what the sparse "fast" form costs
// (1 S-box + 23 full multiplications), the dense 11 x 11 INIT layer disappears, and the whole permutation is ~21 KB of
// code.  CAP_ZERO / OUT4 as in permute_fast.
template <int SBOX_FMA_MASK = 0, int PART_FMA_MASK = 0, bool COLUMN = false, bool CVT_I2F = false, bool COMBINE_ALU = true,
          bool CAP_ZERO = false, bool OUT4 = false>
__device__ __forceinline__ void permute_rounds(uint64_t (&s)[WIDTH]) {
#pragma unroll
  for (int i = 0; i < (CAP_ZERO ? 8 : WIDTH); i++) s[i] = glx::add_canonical(s[i], PMT_RC[i]);
#pragma unroll 1
  for (int r = 0; r < PMT_ROUNDS; r++) {
    if (r < PMT_FULL_HALF || r >= PMT_FULL_HALF + PMT_PARTIAL) {
#pragma unroll
      for (int i = 0; i < 8; i++) s[i] = pow7_mix<SBOX_FMA_MASK>(s[i]);
      if (CAP_ZERO && r == 0) {
#pragma unroll
        for (int i = 8; i < WIDTH; i++) s[i] = PMT_SBOX_RC_CAP[i - 8];
      } else {
#pragma unroll
        for (int i = 8; i < WIDTH; i++) s[i] = pow7_mix<SBOX_FMA_MASK>(s[i]);
      }
    } else {
      s[0] = pow7_mix<PART_FMA_MASK>(s[0]);
    }
    mds_layer_dfma2<COLUMN, CVT_I2F, COMBINE_ALU>(s, &PMT_RC_DM[2 * WIDTH * r], OUT4 && r == PMT_ROUNDS - 1);
  }
}

// Paired form (production): full rounds as in permute_rounds; the 22 partial rounds as 11 PAIRS.  Lanes 1..11 of the
// state between the two rounds of a pair never meet an S-box, so (tools/gen_constants.py::derive_paired)
//     z = A s' + col0(M) x + K,   s' = [sbox(s0), s1..s11],  x = sbox(row0(M) s' + c[0]),  A = M[:,1:] M[1:,:]
// which is 24 + 288 + 24 DFMAs per PAIR of rounds instead of 2 x 288, one conversion of the state to doubles and one
// recombination instead of two.  A's entries are < 2^15 and its row sums < 2^17, so the fp64 sums stay exact (< 2^50).
// The 288 DFMAs of A s' do not depend on the second S-box, whose serial chain they can overlap.
template <int SBOX_FMA_MASK = 0, int PART_FMA_MASK = 0, bool COLUMN = false, bool CVT_I2F = false, bool COMBINE_ALU = true,
          bool CAP_ZERO = false, bool OUT4 = false>
__device__ __forceinline__ void permute_paired(uint64_t (&s)[WIDTH]) {
  const double MAGIC = 4503599627370496.0;  // 2^52
#pragma unroll
  for (int i = 0; i < (CAP_ZERO ? 8 : WIDTH); i++) s[i] = glx::add_canonical(s[i], PMT_RC[i]);
#pragma unroll 1
  for (int half = 0; half < 2; half++) {
#pragma unroll 1
    for (int q = 0; q < PMT_FULL_HALF; q++) {
      const int r = half ? PMT_FULL_HALF + PMT_PARTIAL + q : q;
#pragma unroll
      for (int i = 0; i < 8; i++) s[i] = pow7_mix<SBOX_FMA_MASK>(s[i]);
      if (CAP_ZERO && r == 0) {
#pragma unroll
        for (int i = 8; i < WIDTH; i++) s[i] = PMT_SBOX_RC_CAP[i - 8];
      } else {
#pragma unroll
        for (int i = 8; i < WIDTH; i++) s[i] = pow7_mix<SBOX_FMA_MASK>(s[i]);
      }
      mds_layer_dfma2<COLUMN, CVT_I2F, COMBINE_ALU>(s, &PMT_RC_DM[2 * WIDTH * r], OUT4 && r == PMT_ROUNDS - 1);
    }
    if (half == 0) {
#pragma unroll 1
      for (int pair = 0; pair < PMT_PARTIAL / 2; pair++) {
        const int r = PMT_FULL_HALF + 2 * pair;
        s[0] = pow7_mix<PART_FMA_MASK>(s[0]);
        double dlo[WIDTH], dhi[WIDTH];
#pragma unroll
        for (int i = 0; i < WIDTH; i++) {
          if (CVT_I2F) { dlo[i] = (double)glx::lo32(s[i]); dhi[i] = (double)glx::hi32(s[i]); }
          else {
            dlo[i] = __hiloint2double(0x43300000, (int)glx::lo32(s[i])) - MAGIC;
            dhi[i] = __hiloint2double(0x43300000, (int)glx::hi32(s[i])) - MAGIC;
          }
        }
        // y0 = row0(M) s' + c_{r+1}[0]
        double L0 = PMT_RC_DM[2 * WIDTH * r], H0 = PMT_RC_DM[2 * WIDTH * r + 1];
#pragma unroll
        for (int i = 0; i < WIDTH; i++) {
          const double c = i == 0 ? PMT_MDS_CIRC_D[12] : PMT_MDS_CIRC_D[i];
          L0 = fma(dlo[i], c, L0); H0 = fma(dhi[i], c, H0);
        }
        const uint64_t y0 = COMBINE_ALU ? combine_magic_alu(L0, H0) : combine_magic_fma(L0, H0);
        // z = A s' + K  (independent of the S-box of y0)
        double L[WIDTH], H[WIDTH];
        const double* __restrict__ kd = &PMT_PP_K_DM[2 * WIDTH * pair];
#pragma unroll
        for (int j = 0; j < WIDTH; j++) {
          double l = kd[2 * j], h = kd[2 * j + 1];
#pragma unroll
          for (int i = 0; i < WIDTH; i++) {
            const double c = PMT_PP_A_D[WIDTH * j + i];
            l = fma(dlo[i], c, l); h = fma(dhi[i], c, h);
          }
          L[j] = l; H[j] = h;
        }
        const uint64_t x = pow7_mix<PART_FMA_MASK>(y0);
        double xlo, xhi;
        if (CVT_I2F) { xlo = (double)glx::lo32(x); xhi = (double)glx::hi32(x); }
        else {
          xlo = __hiloint2double(0x43300000, (int)glx::lo32(x)) - MAGIC;
          xhi = __hiloint2double(0x43300000, (int)glx::hi32(x)) - MAGIC;
        }
#pragma unroll
        for (int j = 0; j < WIDTH; j++) {
          const double c = j == 0 ? PMT_MDS_CIRC_D[12] : PMT_MDS_CIRC_D[(WIDTH - j) % WIDTH];   // M[j][0]
          L[j] = fma(xlo, c, L[j]); H[j] = fma(xhi, c, H[j]);
          s[j] = COMBINE_ALU ? combine_magic_alu(L[j], H[j]) : combine_magic_fma(L[j], H[j]);
        }
      }
    }
  }
}

// Paired form with the MDS layers in the frequency domain (poseidon_freq.cuh): 204 fp64 operations per full layer instead
// of 288, 260 per pair of partial rounds instead of 336 -- the same digests (tests/cpp/check_freq.cpp, tests/test_gpu_parity.py).
// FQ_SPLIT: finish the low halves before the high halves are converted (lower register pressure) instead of leaving the
// order to ptxas.
template <int SBOX_FMA_MASK = 0, int PART_FMA_MASK = 0, bool CVT_I2F = true, bool CAP_ZERO = false, bool OUT4 = false,
          int FQ_SPLIT = 1, int COMBINE_MODE = 0>
__device__ __forceinline__ void permute_paired_freq(uint64_t (&s)[WIDTH]) {
  const double MAGIC = 4503599627370496.0;  // 2^52
  const uint32_t zero = PMT_ZERO32;
  auto half_of = [&](uint64_t x, int h) -> double {
    const uint32_t w = h ? glx::hi32(x) : glx::lo32(x);
    return CVT_I2F ? (double)w : __hiloint2double(0x43300000, (int)w) - MAGIC;
  };
#pragma unroll
  for (int i = 0; i < (CAP_ZERO ? 8 : WIDTH); i++) s[i] = glx::add_canonical(s[i], PMT_RC[i]);
#pragma unroll 1
  for (int half = 0; half < 2; half++) {
#pragma unroll 1
    for (int q = 0; q < PMT_FULL_HALF; q++) {
      const int r = half ? PMT_FULL_HALF + PMT_PARTIAL + q : q;
#pragma unroll
      for (int i = 0; i < 8; i++) s[i] = pow7_mix<SBOX_FMA_MASK>(s[i]);
      if (CAP_ZERO && r == 0) {
#pragma unroll
        for (int i = 8; i < WIDTH; i++) s[i] = PMT_SBOX_RC_CAP[i - 8];
      } else {
#pragma unroll
        for (int i = 8; i < WIDTH; i++) s[i] = pow7_mix<SBOX_FMA_MASK>(s[i]);
      }
      if (OUT4 && r == PMT_ROUNDS - 1) {
        mds_layer_dfma2<false, CVT_I2F, true>(s, &PMT_RC_DM[2 * WIDTH * r], true);   // 4 rows: the matrix form is shorter
      } else {
        double x[WIDTH], olo[WIDTH], ohi[WIDTH];
#pragma unroll
        for (int i = 0; i < WIDTH; i++) x[i] = half_of(s[i], 0);
        freq::full_layer_half<2>(x, &PMT_RC_DM[2 * WIDTH * r], olo);
#pragma unroll
        for (int i = 0; i < WIDTH; i++) {
          // fence (see tie()): the high halves are converted after the low halves are done, or ptxas interleaves both
          // halves and needs 200 registers
          if (FQ_SPLIT == 1) x[i] = half_of(tie_hi(s[i], hi_word(olo[i]), zero), 1);
          else x[i] = half_of(s[i], 1);
        }
        freq::full_layer_half<2>(x, &PMT_RC_DM[2 * WIDTH * r + 1], ohi);
#pragma unroll
        for (int i = 0; i < WIDTH; i++) s[i] = (COMBINE_MODE == 1 || COMBINE_MODE == 2) ? combine_magic_fma(olo[i], ohi[i]) : combine_magic_alu(olo[i], ohi[i]);
      }
    }
    if (half == 0) {
#pragma unroll 1
      for (int pair = 0; pair < PMT_PARTIAL / 2; pair++) {
        const int r = PMT_FULL_HALF + 2 * pair;
        s[0] = pow7_mix<PART_FMA_MASK>(s[0]);
        double x[WIDTH], ylo[WIDTH], yhi[WIDTH], l0lo, l0hi, x0lo, x0hi;
#pragma unroll
        for (int i = 0; i < WIDTH; i++) x[i] = half_of(s[i], 0);
        freq::pair_half_begin(x, PMT_RC_DM[2 * WIDTH * r], l0lo, ylo, x0lo);
#pragma unroll
        for (int i = 0; i < WIDTH; i++) {
          if (FQ_SPLIT == 1) x[i] = half_of(tie_hi(s[i], hi_word(ylo[i]), zero), 1);
          else x[i] = half_of(s[i], 1);
        }
        freq::pair_half_begin(x, PMT_RC_DM[2 * WIDTH * r + 1], l0hi, yhi, x0hi);
        const uint64_t xs = pow7_mix<PART_FMA_MASK>(combine_magic_alu(l0lo, l0hi));
        freq::pair_half_end<2>(ylo, half_of(xs, 0), x0lo, l0lo, &PMT_FQ_KPAIR_DM[2 * WIDTH * pair]);
        freq::pair_half_end<2>(yhi, half_of(xs, 1), x0hi, l0hi, &PMT_FQ_KPAIR_DM[2 * WIDTH * pair + 1]);
#pragma unroll
        for (int i = 0; i < WIDTH; i++) s[i] = (COMBINE_MODE == 1 || COMBINE_MODE == 3) ? combine_magic_fma(ylo[i], yhi[i]) : combine_magic_alu(ylo[i], yhi[i]);
      }
    }
  }
}

// six carry-free column sums of sum_t x[t] * K[t], K as 22/22/20-bit limbs kl[3 t + j]
struct DotAcc { uint64_t t00, t01, t02, t10, t11, t12; };
template <int N>
__device__ __forceinline__ void dot_acc(DotAcc& a, const uint64_t* x, const uint32_t* __restrict__ kl) {
#pragma unroll
  for (int t = 0; t < N; t++) {
    const uint32_t a0 = glx::lo32(x[t]), a1 = glx::hi32(x[t]);
    const uint32_t b0 = kl[3 * t], b1 = kl[3 * t + 1], b2 = kl[3 * t + 2];
    a.t00 = glx::mad_wide(a0, b0, a.t00); a.t01 = glx::mad_wide(a0, b1, a.t01); a.t02 = glx::mad_wide(a0, b2, a.t02);
    a.t10 = glx::mad_wide(a1, b0, a.t10); a.t11 = glx::mad_wide(a1, b1, a.t11); a.t12 = glx::mad_wide(a1, b2, a.t12);
  }
}
template <bool ALU>
__device__ __forceinline__ uint64_t dot_finish(const DotAcc& a) {
  // V = G0 + 2^32 G1,  G0 = t00 + 2^22 t01 + 2^44 t02 (< 2^103), G1 likewise;  2^32 G1 = 2^32 g_lo - g_hi (mod p)
  glx::u128 g0 = (glx::u128)a.t00 + ((glx::u128)a.t01 << 22) + ((glx::u128)a.t02 << 44);
  glx::u128 g1 = (glx::u128)a.t10 + ((glx::u128)a.t11 << 22) + ((glx::u128)a.t12 << 44);
  const uint64_t g_lo = (uint64_t)g1, g_hi = (uint64_t)(g1 >> 64);
  glx::u128 v = g0 + ((glx::u128)g_lo << 32) + (((glx::u128)glx::P << 40) - g_hi);
  return glx::reduce128<ALU>(v);
}

// SBOX_FMA_MASK / PART_FMA_MASK: pow7_mix masks of the full / partial rounds; MULADD_ALU, DOT_ALU: reduction pipe of the
// partial rounds' multiply-adds and dot products.  CAP_ZERO / OUT4 as in permute_fast.
template <int SBOX_FMA_MASK = 0, int PART_FMA_MASK = 0, bool MULADD_ALU = true, bool DOT_ALU = false, bool CVT_I2F = false,
          bool COMBINE_ALU = true, bool CAP_ZERO = false, bool OUT4 = false, int PIPE = 0, int PPIPE = 0>
__device__ __forceinline__ void permute_fused(uint64_t (&s)[WIDTH]) {
#pragma unroll
  for (int i = 0; i < (CAP_ZERO ? 8 : WIDTH); i++) s[i] = glx::add_canonical(s[i], PMT_RC[i]);
#pragma unroll 1
  for (int half = 0; half < 2; half++) {
#pragma unroll 1
    for (int r = 0; r < PMT_FULL_HALF; r++)
      full_round_fused<SBOX_FMA_MASK, CVT_I2F, COMBINE_ALU, PIPE>(s, &PMT_RC_AFTER_FULL_DM[2 * WIDTH * (PMT_FULL_HALF * half + r)],
                                                            CAP_ZERO && half == 0 && r == 0,
                                                            OUT4 && half == 1 && r == PMT_FULL_HALF - 1);
    if (half == 0) {
      {  // dense INIT matrix on lanes 1..11 (lane 0 passes through)
        uint64_t y[WIDTH];
#pragma unroll 1
        for (int a = 1; a < WIDTH; a++) {
          DotAcc acc = {0, 0, 0, 0, 0, 0};
          dot_acc<WIDTH - 1>(acc, &s[1], &PMT_FP_INIT_L11[3 * (WIDTH - 1) * (a - 1)]);
          const uint64_t v = dot_finish<DOT_ALU>(acc);
#pragma unroll
          for (int i = 1; i < WIDTH; i++) if (i == a) y[i] = v;   // static indexing keeps y[] in registers
        }
#pragma unroll
        for (int i = 1; i < WIDTH; i++) s[i] = y[i];
      }
#pragma unroll 1
      for (int k = 0; k < PMT_PARTIAL; k++) {
        DotAcc acc = {0, 0, 0, 0, 0, 0};
        const uint32_t* __restrict__ wl = &PMT_FP_W_HAT_L11[3 * (WIDTH - 1) * k];
        uint64_t x0;
        if (PPIPE == 0) {
          dot_acc<WIDTH - 1>(acc, &s[1], wl);   // independent of the S-box below
          x0 = glx::add_canonical(pow7_mix<PART_FMA_MASK>(s[0]), PMT_FP_POST_RC[k]);
        } else {
          // the S-box's four dependent multiplications are staged against quarters of the dot product
          const uint32_t zero = PMT_ZERO32;
          const uint64_t x = s[0];
          dot_acc<3>(acc, &s[1], wl);
          const uint64_t x2 = glx::sqr<!(PART_FMA_MASK & 1)>(x);
          dot_acc<3>(acc, &s[4], wl + 9);
          const uint64_t x4 = glx::sqr<!(PART_FMA_MASK & 2)>(tie(x2, glx::hi32(acc.t12), zero));
          dot_acc<3>(acc, &s[7], wl + 18);
          const uint64_t x3 = glx::mul<!(PART_FMA_MASK & 4)>(x, tie(x2, glx::hi32(acc.t12), zero));
          dot_acc<2>(acc, &s[10], wl + 27);
          x0 = glx::add_canonical(glx::mul<!(PART_FMA_MASK & 8)>(x3, tie(x4, glx::hi32(acc.t12), zero)), PMT_FP_POST_RC[k]);
        }
        acc.t00 = glx::mad_wide(glx::lo32(x0), (uint32_t)PMT_FP_M00, acc.t00);
        acc.t10 = glx::mad_wide(glx::hi32(x0), (uint32_t)PMT_FP_M00, acc.t10);
        const uint64_t d = dot_finish<DOT_ALU>(acc);
#pragma unroll
        for (int i = 1; i < WIDTH; i++) s[i] = glx::mul_add<MULADD_ALU>(x0, PMT_FP_V[(WIDTH - 1) * k + (i - 1)], s[i]);
        s[0] = d;
      }
#pragma unroll
      for (int i = 0; i < WIDTH; i++) s[i] = glx::add_canonical(s[i], PMT_RC[WIDTH * (PMT_FULL_HALF + PMT_PARTIAL) + i]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Cooperative form for the latency-bound levels: 16 lanes share ONE state (lane g < 12 holds element g, lanes 12..15
// idle).  A lone warp needs ~58 us for a thread-per-state permutation (25 k dependent-ish instructions); here the 12
// S-boxes of a round run side by side and the MDS row of every lane is 11 pairs of warp shuffles + 24 IMAD.WIDE, so a
// permutation is ~3.8 k instructions per warp and ~5 us.  Specification form (30 x constants, S-box, MDS): in the
// partial rounds every lane computes the S-box (SIMT) and only lane 0 keeps it.
// rc: the 372-entry table PMT_RC staged in SHARED memory (lane-indexed reads from constant memory would serialise).
// All 32 lanes of the warp must call this together.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t permute_coop(uint64_t v, const uint64_t* __restrict__ rc, unsigned g,
                                                 unsigned group_base_lane) {
  constexpr uint32_t CIRC[WIDTH] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
  const unsigned gg = g < WIDTH ? g : 0;
  const uint32_t c0 = g == 0 ? 25u : 17u;  // CIRC[0] + DIAG[0] on lane 0
  v = glx::add_canonical(v, rc[gg]);
#pragma unroll 1
  for (int r = 0; r < PMT_ROUNDS; r++) {
    const bool full = r < PMT_FULL_HALF || r >= PMT_FULL_HALF + PMT_PARTIAL;
    const uint64_t p = glx::pow7<true>(v);
    v = (full || g == 0) ? p : v;
    const uint32_t lo = glx::lo32(v), hi = glx::hi32(v);
    uint64_t L = glx::mad_wide(lo, c0, rc[WIDTH * (r + 1) + gg]);   // constants < 2^64 - 2^48, sums < 2^42: no overflow
    uint64_t H = (uint64_t)hi * c0;
#pragma unroll
    for (int i = 1; i < WIDTH; i++) {
      const unsigned idx = gg + i;
      const unsigned src = group_base_lane + (idx >= WIDTH ? idx - WIDTH : idx);
      L = glx::mad_wide(__shfl_sync(0xffffffffu, lo, src), CIRC[i], L);
      H = glx::mad_wide(__shfl_sync(0xffffffffu, hi, src), CIRC[i], H);
    }
    v = glx::combine_halves(L, H);
  }
  return v;
}

}  // namespace poseidonx
