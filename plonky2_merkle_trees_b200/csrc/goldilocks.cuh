// goldilocks.cuh -- Goldilocks field (p = 2^64 - 2^32 + 1) arithmetic for sm_100a integer pipes.
//
// Replaces [UPSTREAM plonky2 field/src/goldilocks_field.rs] (GoldilocksField add / mul / reduce128 /
// to_canonical_u64), the arithmetic under PoseidonHash at /root/reference/src/simple_merkle_tree/
// simple_merkle_tree.rs:23,33 and /root/reference/src/mmr/merkle_mountain_ranges.rs:96,111.
//
// Representation: a felt is any u64 congruent to the value mod p ("non-canonical allowed", as upstream);
// gl::canonical() is applied once, when a digest leaves the permutation.
//
// B200 mapping (measured): a 64x64 multiply is 4 IMAD.WIDE.U32 (fma-heavy pipe, 4 cycles each per SM
// sub-partition) + carry glue and the reduction on the ALU pipe (IADD3[.X], 2 cycles each).  The reduction uses
// 2^64 = 2^32 - 1 and 2^96 = -1 (mod p).  Only the forms the product kernels use live here; the other reductions that were
// measured (PTX carry chains, IMAD.HI folds, IMAD.WIDE recombination) are in tools/experimental/goldilocks_variants.cuh.
#pragma once
#include <stdint.h>

namespace gl {

typedef unsigned __int128 u128;
static constexpr uint64_t P = 0xFFFFFFFF00000001ull;
static constexpr uint32_t EPS = 0xFFFFFFFFu;  // 2^64 mod p

__device__ __forceinline__ uint32_t lo32(uint64_t x) { return (uint32_t)x; }
__device__ __forceinline__ uint32_t hi32(uint64_t x) { return (uint32_t)(x >> 32); }
__device__ __forceinline__ uint64_t pack(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

__device__ __forceinline__ uint64_t canonical(uint64_t x) { return x >= P ? x - P : x; }

// acc + a * b as ONE chained IMAD.WIDE.U32 Rd, Ra, Rb, Rd.  Written as the mad.lo.cc / madc.hi pair because that is the only
// spelling ptxas 12.9 keeps as an accumulate (plain C is re-associated into independent IMAD.WIDE + IADD3 add trees).  The
// 64-bit sum must not overflow (callers keep it < 2^63).
__device__ __forceinline__ uint64_t mad_wide(uint32_t a, uint32_t b, uint64_t acc) {
  uint32_t lo = (uint32_t)acc, hi = (uint32_t)(acc >> 32);
  asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
  return ((uint64_t)hi << 32) | lo;
}

// Signed two-word form of the reduction, ANY 128-bit input (w3:w2:w1:w0):
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

__device__ __forceinline__ uint64_t reduce128(u128 v) {
  const uint64_t lo = (uint64_t)v, hi = (uint64_t)(v >> 64);
  return reduce128_c(lo32(lo), hi32(lo), lo32(hi), hi32(hi));
}

__device__ __forceinline__ uint64_t mul(uint64_t a, uint64_t b) { return reduce128((u128)a * b); }

// l + 2^32 h (mod p) for any l and h < 2^63 (the column sums of the integer MDS rows of the 16-lane form): three words, one reduction
__device__ __forceinline__ uint64_t combine_sums(uint64_t l, uint64_t h) {
  const uint64_t mid = (uint64_t)hi32(l) + lo32(h);                   // < 2^33
  const uint32_t w2 = hi32(h) + hi32(mid);                            // < 2^32: h < 2^63
  return reduce128_c(lo32(l), lo32(mid), w2, 0u);
}

// a^2 with THREE IMAD.WIDE instead of four: a0^2 + 2^33 a0 a1 + 2^64 a1^2.  nvcc's (u128)a * a computes a0 a1 twice (the
// second time as the accumulate that doubles it) and needs an IMAD.X + IMAD.MOV to carry the 65th bit into a1^2; here the
// cross product is doubled by an add and two funnel shifts, and (2m >> 32) + the carry of the middle word ride on the
// addend / carry-in of the last multiply: 3 IMAD.WIDE + 1 IMAD.IADD + 11 alu instead of 4 IMAD.WIDE + 2 IMAD + 9 alu.
// Exact: the sum is a^2 < 2^128, word by word (tools/check_reduce.c).
__device__ __forceinline__ uint64_t sqr(uint64_t a) {
  const uint32_t a0 = lo32(a), a1 = hi32(a);
  uint32_t p0, p1, m0, m1, w1, w2, w3;
  asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %2;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(p0), "=r"(p1) : "r"(a0));
  asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(m0), "=r"(m1) : "r"(a0), "r"(a1));
  const uint32_t d0 = m0 << 1, d1 = __funnelshift_l(m0, m1, 1), d2 = m1 >> 31;   // 2m = d0 + 2^32 d1 + 2^64 d2
  asm("add.cc.u32 %0, %3, %4;\n\tmadc.lo.cc.u32 %1, %5, %5, %6;\n\tmadc.hi.u32 %2, %5, %5, %7;"
      : "=r"(w1), "=r"(w2), "=r"(w3) : "r"(p1), "r"(d0), "r"(a1), "r"(d1), "r"(d2));
  return reduce128_c(p0, w1, w2, w3);
}

// a + c where c is canonical (< p); result any u64 congruent to the sum
__device__ __forceinline__ uint64_t add_canonical(uint64_t a, uint64_t c) {
  uint64_t r = a + c;
  if (r < a) r += EPS;  // a + c < 2^64 + p, so the wrapped value is < p and cannot wrap again
  return r;
}

// the S-box x^7: two squarings and two multiplications, 14 IMAD.WIDE
__device__ __forceinline__ uint64_t pow7(uint64_t x) {
  const uint64_t x2 = sqr(x), x4 = sqr(x2), x3 = mul(x, x2);
  return mul(x3, x4);
}

}  // namespace gl
