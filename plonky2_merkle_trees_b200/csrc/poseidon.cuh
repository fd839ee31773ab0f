// poseidon.cuh -- width-12 Poseidon permutation over Goldilocks, one state per thread, in registers (production form).
//
// Replaces [UPSTREAM plonky2 hash/poseidon.rs Poseidon::poseidon, hash/poseidon_goldilocks.rs, hash/hashing.rs
// compress / hash_n_to_m_no_pad], i.e. what PoseidonHash::{two_to_one, hash_or_noop} execute for
// /root/reference/src/simple_merkle_tree/simple_merkle_tree.rs:23,33,45 and
// /root/reference/src/mmr/merkle_mountain_ranges.rs:96,111,125.
//
// sm_100a mapping (no tensor cores: this is not a contraction):
//   * state = 12 x u64 = 24 registers per thread; 4 full rounds, 11 PAIRS of partial rounds, 4 full rounds, each group one
//     rolled loop so the whole permutation is 31 KB of SASS (inside the instruction cache).
//   * S-box x^7 = 2 squarings + 2 multiplications on the integer pipes (gl::pow7: 14 IMAD.WIDE + ALU-only reductions).
//   * every MDS layer runs on the fp64 pipe as a length-12 cyclic convolution in the frequency domain
//     (poseidon_freq.cuh): lanes are split in 32-bit halves converted with I2F, accumulators carry 2^52 + constant so the
//     mantissa IS the integer, and the two halves are recombined mod p on the ALU pipe (combine_magic).
//   * partial rounds in pairs (tools/gen_constants.py::derive_paired): lanes 1..11 of the state between the two rounds of
//     a pair never meet an S-box, so z = A s' + col0(M) x + K with x = sbox(row0(M) s' + c): one conversion and one
//     recombination of the state per pair instead of two.
// The other forms that were built and measured (specification form, sparse "fast" partial rounds, IMAD.WIDE / limb / plain
// DFMA MDS layers, the fused S-box + column form, the fp64 tensor-pipe quad layout) live in tools/experimental/ and are
// compiled only by the A/B harness tools/ab_level.cu.
#pragma once
#include "goldilocks.cuh"
#include "poseidon_constants.cuh"
#include "poseidon_freq.cuh"

namespace poseidon {

static constexpr int WIDTH = 12;

// 2^52 + l, 2^52 + h (l, h < 2^52) -> u64 congruent to l + 2^32 h, ALU pipe only, plain C (signed two-word form, like
// gl::reduce128_c): V = (l0 - h1) + 2^32 (l1 + h0 + h1); the exponent bits 0x43300000 of both high words are removed by
// constants folded into the 3-input adds, so there is no masking.  Exact for every l, h < 2^52 (tools/check_combine.c).
__device__ __forceinline__ uint64_t combine_magic(double L, double H) {
  const uint64_t lb = (uint64_t)__double_as_longlong(L), hb = (uint64_t)__double_as_longlong(H);
  const int64_t lo = (int64_t)(uint64_t)gl::lo32(lb) - (int64_t)(uint64_t)gl::hi32(hb) + 0x43300000ll;
  const int64_t hi = (int64_t)(uint64_t)gl::hi32(lb) + (int64_t)(uint64_t)gl::lo32(hb) + (int64_t)(uint64_t)gl::hi32(hb) +
                     (lo >> 32) - 0x86600000ll;
  const int64_t n = hi >> 32;
  return gl::pack((uint32_t)lo, (uint32_t)hi) + ((uint64_t)n << 32) - (uint64_t)n;
}

// the 32-bit half h of x as a double: I2F.F64.U32 on the (otherwise idle) conversion pipe
__device__ __forceinline__ double half_of(uint64_t x, int h) { return (double)(h ? gl::hi32(x) : gl::lo32(x)); }

// The permutation.  Output lanes are NOT canonicalised (any u64 congruent to the value).  Same digests as the specification
// form: tests/cpp/check_freq.cpp (host, 200 000 states + every layer at all corners of the input cube) and
// tests/test_gpu_parity.py (device, against the CPU restatement).
__device__ __forceinline__ void permute(uint64_t (&s)[WIDTH]) {
#pragma unroll
  for (int i = 0; i < WIDTH; i++) s[i] = gl::add_canonical(s[i], PMT_RC[i]);
#pragma unroll 1
  for (int half = 0; half < 2; half++) {
#pragma unroll 1
    for (int q = 0; q < PMT_FULL_HALF; q++) {
      const int r = half ? PMT_FULL_HALF + PMT_PARTIAL + q : q;
#pragma unroll
      for (int i = 0; i < WIDTH; i++) s[i] = gl::pow7(s[i]);
      double x[WIDTH], olo[WIDTH], ohi[WIDTH];
#pragma unroll
      for (int i = 0; i < WIDTH; i++) x[i] = half_of(s[i], 0);
      freq::full_layer_half<2>(x, &PMT_RC_DM[2 * WIDTH * r], olo);
#pragma unroll
      for (int i = 0; i < WIDTH; i++) x[i] = half_of(s[i], 1);
      freq::full_layer_half<2>(x, &PMT_RC_DM[2 * WIDTH * r + 1], ohi);
#pragma unroll
      for (int i = 0; i < WIDTH; i++) s[i] = combine_magic(olo[i], ohi[i]);
    }
    if (half == 0) {
#pragma unroll 1
      for (int pair = 0; pair < PMT_PARTIAL / 2; pair++) {
        const int r = PMT_FULL_HALF + 2 * pair;
        s[0] = gl::pow7(s[0]);
        double x[WIDTH], ylo[WIDTH], yhi[WIDTH], l0lo, l0hi, x0lo, x0hi;
#pragma unroll
        for (int i = 0; i < WIDTH; i++) x[i] = half_of(s[i], 0);
        freq::pair_half_begin(x, PMT_RC_DM[2 * WIDTH * r], l0lo, ylo, x0lo);
#pragma unroll
        for (int i = 0; i < WIDTH; i++) x[i] = half_of(s[i], 1);
        freq::pair_half_begin(x, PMT_RC_DM[2 * WIDTH * r + 1], l0hi, yhi, x0hi);
        const uint64_t xs = gl::pow7(combine_magic(l0lo, l0hi));
        freq::pair_half_end<2>(ylo, half_of(xs, 0), x0lo, l0lo, &PMT_FQ_KPAIR_DM[2 * WIDTH * pair]);
        freq::pair_half_end<2>(yhi, half_of(xs, 1), x0hi, l0hi, &PMT_FQ_KPAIR_DM[2 * WIDTH * pair + 1]);
#pragma unroll
        for (int i = 0; i < WIDTH; i++) s[i] = combine_magic(ylo[i], yhi[i]);
      }
    }
  }
}

}  // namespace poseidon
