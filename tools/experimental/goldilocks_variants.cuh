// goldilocks_variants.cuh -- EXPERIMENTAL (tools/ab_level.cu only; the product uses csrc/goldilocks.cuh): every reduction /
// multiplication form that was measured in round 1, in namespace glx.
// goldilocks.cuh -- Goldilocks field (p = 2^64 - 2^32 + 1) arithmetic for sm_100a integer pipes.
//
// Replaces [UPSTREAM plonky2 field/src/goldilocks_field.rs] (GoldilocksField add / mul / reduce128 /
// to_canonical_u64), the arithmetic under PoseidonHash at /root/reference/src/simple_merkle_tree/
// simple_merkle_tree.rs:23,33 and /root/reference/src/mmr/merkle_mountain_ranges.rs:96,111.
//
// Representation: a felt is any u64 congruent to the value mod p ("non-canonical allowed", as upstream);
// gl::canonical() is applied once, when a digest leaves the permutation.
//
// B200 mapping: a 64x64 multiply is 4 IMAD.WIDE.U32 (fma pipe) + carry glue on the ALU pipe (IADD3[.X]); the
// reduction uses 2^64 = 2^32 - 1 and 2^96 = -1 (mod p), so it is three 32-bit add/sub chains plus two
// IMAD.HI folds -- 16 SASS instructions per general multiply (checked with cuobjdump).
#pragma once
#include <stdint.h>

namespace glx {

typedef unsigned __int128 u128;
static constexpr uint64_t P = 0xFFFFFFFF00000001ull;
static constexpr uint32_t EPS = 0xFFFFFFFFu;  // 2^64 mod p

__device__ __forceinline__ uint32_t lo32(uint64_t x) { return (uint32_t)x; }
__device__ __forceinline__ uint32_t hi32(uint64_t x) { return (uint32_t)(x >> 32); }
__device__ __forceinline__ uint64_t pack(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// acc + a * b as ONE chained IMAD.WIDE.U32 Rd, Ra, Rb, Rd.  Written as the mad.lo.cc / madc.hi pair because that is
// the only spelling ptxas 12.9 keeps as an accumulate: plain C (or mad.wide.u32) is re-associated into independent
// IMAD.WIDE + 3-input IADD3/IADD3.X add trees and power-of-two coefficients into shifts -- 2x the instructions, all on
// the ALU pipe, which is the scarcer pipe in this kernel.  The 64-bit sum must not overflow (callers keep it < 2^63).
__device__ __forceinline__ uint64_t mad_wide(uint32_t a, uint32_t b, uint64_t acc) {
  uint32_t lo = (uint32_t)acc, hi = (uint32_t)(acc >> 32);
  asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
  return ((uint64_t)hi << 32) | lo;
}

__device__ __forceinline__ uint64_t canonical(uint64_t x) { return x >= P ? x - P : x; }

// (w3:w2:w1:w0) mod p, any 128-bit input whose top word is < 2^32 - 1 (true for every product of two u64).
//   V = w0 + 2^32 w1 + (2^32 - 1) w2 - w3  =  [(w1:w0) - (w3:w2)] + 2^32 (w3:w2)      (mod p)
__device__ __forceinline__ uint64_t reduce128(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
  uint32_t c, t;
  asm("sub.cc.u32 %0, %0, %3;\n\tsubc.cc.u32 %1, %1, %4;\n\tsubc.u32 %2, 0, 0;"
      : "+r"(w0), "+r"(w1), "=r"(c) : "r"(w2), "r"(w3));
  asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(w1), "+r"(c) : "r"(w2), "r"(w3));
  // value = (w1:w0) + 2^64 c with 0 <= c <= w3 + 1; fold c * (2^32 - 1), then the (rare) carry once more
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, 0, 0;"
      : "+r"(w0), "+r"(w1), "=r"(t) : "r"(c), "r"(EPS));
  asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(w0), "+r"(w1) : "r"(t), "r"(EPS));
  return pack(w0, w1);
}

// Same reduction with the two c*(2^32-1) folds done as (c << 32) - c on the ALU pipe (no IMAD.HI, which issues at half
// rate on B200): 13 IADD3-class instructions, zero fma-pipe slots.  Used where the fma pipe is the bottleneck.
__device__ __forceinline__ uint64_t reduce128_alu(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
  uint32_t c, c2;
  asm("sub.cc.u32 %0, %0, %3;\n\tsubc.cc.u32 %1, %1, %4;\n\tsubc.u32 %2, 0, 0;"
      : "+r"(w0), "+r"(w1), "=r"(c) : "r"(w2), "r"(w3));
  asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(w1), "+r"(c) : "r"(w2), "r"(w3));
  // (w1:w0) + (c << 32) - c ; c2 = carry - borrow.  A borrow means (w1:w0) < c < 2^32, so the wrapped high word is
  // 0xffffffff and adding c >= 1 to it always carries: c2 is 0 or 1, never -1.
  asm("sub.cc.u32 %0, %0, %3;\n\tsubc.cc.u32 %1, %1, 0;\n\tsubc.u32 %2, 0, 0;" : "+r"(w0), "+r"(w1), "=r"(c2) : "r"(c));
  asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, 0;" : "+r"(w1), "+r"(c2) : "r"(c));
  // second fold: + (c2 << 32) - c2.  If c2 = 1 the wrapped value is < 2^64 - 2^32, so this cannot wrap again.
  asm("sub.cc.u32 %0, %0, %2;\n\tsubc.u32 %1, %1, 0;" : "+r"(w0), "+r"(w1) : "r"(c2));
  w1 += c2;
  return pack(w0, w1);
}

// Signed two-word form of the same reduction, ANY 128-bit input:
//   V = (w0 - w2 - w3) + 2^32 (w1 + w2)   (mod p),   lo = w0 - w2 - w3 in (-2^33, 2^32),  hi = w1 + w2 + (lo >> 32)
// hi overflows 32 bits by n in {-1, 0, 1}; n 2^64 = n (2^32 - 1) = (n << 32) - n is added in 64-bit arithmetic, which
// cannot wrap again (checked exhaustively on the host against __int128 % p, tools/check_reduce.c).  Written in plain C so
// that ptxas uses 3-input IADD3 / IADD3.X with two carry predicates: 7 ALU instructions instead of 13.
__device__ __forceinline__ uint64_t reduce128_c(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
  const int64_t lo = (int64_t)(uint64_t)w0 - (int64_t)(uint64_t)w2 - (int64_t)(uint64_t)w3;
  const int64_t hi = (int64_t)(uint64_t)w1 + (int64_t)(uint64_t)w2 + (lo >> 32);
  const int64_t n = hi >> 32;
  return pack((uint32_t)lo, (uint32_t)hi) + ((uint64_t)n << 32) - (uint64_t)n;
}

#ifndef PMT_REDUCE_C
#define PMT_REDUCE_C 1   // 1: the ALU-only reduction is reduce128_c (3-input adds), 0: the hand-written PTX carry chain
#endif
template <bool ALU = false>
__device__ __forceinline__ uint64_t reduce128(u128 v) {
  uint64_t lo = (uint64_t)v, hi = (uint64_t)(v >> 64);
  if (ALU && PMT_REDUCE_C) return reduce128_c(lo32(lo), hi32(lo), lo32(hi), hi32(hi));
  return ALU ? reduce128_alu(lo32(lo), hi32(lo), lo32(hi), hi32(hi)) : reduce128(lo32(lo), hi32(lo), lo32(hi), hi32(hi));
}

template <bool ALU = false>
__device__ __forceinline__ uint64_t mul(uint64_t a, uint64_t b) { return reduce128<ALU>((u128)a * b); }
// a^2 with THREE IMAD.WIDE instead of four: a0^2 + 2^33 a0 a1 + 2^64 a1^2.  nvcc's (u128)a * a computes a0 a1 twice (the
// second time as the accumulate that doubles it) and needs an IMAD.X + IMAD.MOV to carry the 65th bit into a1^2; here the
// cross product is doubled by an add and two funnel shifts, and (2m >> 32) + the carry of the middle word ride on the
// addend / carry-in of the last multiply: 3 IMAD.WIDE + 1 IMAD.IADD + 11 alu instead of 4 IMAD.WIDE + 2 IMAD + 9 alu --
// 6 cycles less on the fma-heavy pipe (the busier one in k_level) for 4 more on the alu pipe.  Exact: the
// sum is a^2 < 2^128, word by word.  PMT_SQR3 selects it (ALU reduction only).
#ifndef PMT_SQR3
#define PMT_SQR3 1
#endif
__device__ __forceinline__ uint64_t sqr3(uint64_t a) {
  const uint32_t a0 = lo32(a), a1 = hi32(a);
  uint32_t p0, p1, m0, m1, w1, w2, w3;
  asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %2;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(p0), "=r"(p1) : "r"(a0));
  asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(m0), "=r"(m1) : "r"(a0), "r"(a1));
#if PMT_SQR3 == 2   // A/B: the low word doubled by a funnel shift (alu pipe) instead of the IMAD.IADD ptxas picks for m0 << 1
  const uint32_t d0 = __funnelshift_l(0u, m0, 1), d1 = __funnelshift_l(m0, m1, 1), d2 = m1 >> 31;
#else
  const uint32_t d0 = m0 << 1, d1 = __funnelshift_l(m0, m1, 1), d2 = m1 >> 31;   // 2m = d0 + 2^32 d1 + 2^64 d2
#endif
  asm("add.cc.u32 %0, %3, %4;\n\tmadc.lo.cc.u32 %1, %5, %5, %6;\n\tmadc.hi.u32 %2, %5, %5, %7;"
      : "=r"(w1), "=r"(w2), "=r"(w3) : "r"(p1), "r"(d0), "r"(a1), "r"(d1), "r"(d2));
  return reduce128_c(p0, w1, w2, w3);
}

template <bool ALU = false>
__device__ __forceinline__ uint64_t sqr(uint64_t a) {
  if (ALU && PMT_REDUCE_C && PMT_SQR3) return sqr3(a);
  return reduce128<ALU>((u128)a * a);
}

// a * b + c  (c any u64): product <= (2^64-1)^2, plus c still fits 128 bits
template <bool ALU = false>
__device__ __forceinline__ uint64_t mul_add(uint64_t a, uint64_t b, uint64_t c) { return reduce128<ALU>((u128)a * b + c); }

// a + c where c is canonical (< p); result any u64 congruent to the sum
__device__ __forceinline__ uint64_t add_canonical(uint64_t a, uint64_t c) {
  uint64_t r = a + c;
  if (r < a) r += EPS;  // a + c < 2^64 + p, so the wrapped value is < p and cannot wrap again
  return r;
}

// a + b for arbitrary u64 operands
__device__ __forceinline__ uint64_t add(uint64_t a, uint64_t b) { return add_canonical(a, canonical(b)); }

template <bool ALU = false>
__device__ __forceinline__ uint64_t pow7(uint64_t x) {
  uint64_t x2 = sqr<ALU>(x), x4 = sqr<ALU>(x2), x3 = mul<ALU>(x, x2);
  return mul<ALU>(x3, x4);
}

// value = lo + 2^32 * hi  ->  u64 congruent mod p.   hi = hh * 2^32 + hl:
//     lo + 2^32 hl + 2^64 hh  =  [lo + (2^32 - 1) hh] + 2^32 hl          (needs lo + (2^32 - 1) hh < 2^64)
// 2 IMAD.WIDE + IADD3 + IADD3.X: the carry of the high-word add is folded back as carry * (2^32 - 1); the wrapped
// high word is then tiny, so the fold cannot carry again.
__device__ __forceinline__ uint64_t combine_halves(uint64_t lo, uint64_t hi) {
  uint64_t a = mad_wide(hi32(hi), EPS, lo);
  uint32_t a0 = lo32(a), a1 = hi32(a), c;
  asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, 0, 0;" : "+r"(a1), "=r"(c) : "r"(lo32(hi)));
  return mad_wide(c, EPS, pack(a0, a1));
}

}  // namespace glx
