// host check of gl::reduce128_c (plonky2_merkle_trees_b200/csrc/goldilocks.cuh) against __int128 % p:
// gcc -O2 -o check_reduce tools/check_reduce.c && ./check_reduce
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
typedef unsigned __int128 u128;
static const uint64_t P = 0xFFFFFFFF00000001ull;
static uint64_t red_v3(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
  int64_t lo = (int64_t)(uint64_t)w0 - (int64_t)(uint64_t)w2 - (int64_t)(uint64_t)w3;
  int64_t hi = (int64_t)(uint64_t)w1 + (int64_t)(uint64_t)w2 + (lo >> 32);
  uint32_t r0 = (uint32_t)lo, r1 = (uint32_t)hi;
  int64_t n = hi >> 32;
  uint64_t r = ((uint64_t)r1 << 32) | r0;
  return r + ((uint64_t)n << 32) - (uint64_t)n;
}
static uint64_t rnd() { static uint64_t s = 88172645463325252ull; s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
int main() {
  uint64_t edge[] = {0, 1, 2, 0xFFFFFFFFull, 0x100000000ull, 0xFFFFFFFF00000000ull, P - 1, P, P + 1, ~0ull, ~0ull - 1, 0x8000000000000000ull, 0xFFFFFFFEFFFFFFFFull};
  int ne = sizeof edge / sizeof *edge; long bad = 0, cnt = 0;
  for (long it = 0; it < 40000000; it++) {
    uint64_t a, b;
    if (it < ne * ne) { a = edge[it / ne]; b = edge[it % ne]; } else { a = rnd(); b = rnd(); if (it % 7 == 0) a |= 0xFFFFFFFF00000000ull; if (it % 11 == 0) b = ~0ull - (b & 0xFFFF); if (it % 13 == 0) a &= 0xFFFFFFFFull; }
    u128 pr = (u128)a * b;
    uint64_t l = (uint64_t)pr, h = (uint64_t)(pr >> 64);
    uint64_t got = red_v3((uint32_t)l, (uint32_t)(l >> 32), (uint32_t)h, (uint32_t)(h >> 32));
    uint64_t want = (uint64_t)(pr % P);
    if (got % P != want) { if (bad < 5) printf("BAD a=%016lx b=%016lx got=%016lx want=%016lx\n", a, b, got, want); bad++; }
    cnt++;
  }
  // arbitrary 128-bit inputs with w3 <= 0xFFFFFFFE too (mul_add: a*b + c)
  for (long it = 0; it < 20000000; it++) {
    uint64_t a = rnd(), b = rnd(), c = rnd(); if (it % 5 == 0) { a = ~0ull; b = ~0ull; c = ~0ull - (it & 255); }
    u128 pr = (u128)a * b + c;
    uint64_t l = (uint64_t)pr, h = (uint64_t)(pr >> 64);
    uint64_t got = red_v3((uint32_t)l, (uint32_t)(l >> 32), (uint32_t)h, (uint32_t)(h >> 32));
    if (got % P != (uint64_t)(pr % P)) { if (bad < 5) printf("BAD2\n"); bad++; }
    cnt++;
  }
  // fully arbitrary 128-bit values
  for (long it = 0; it < 20000000; it++) {
    uint64_t l = rnd(), h = rnd(); if (it % 3 == 0) h |= 0xFFFFFFFF00000000ull; if (it % 5 == 0) l = 0; if (it % 7 == 0) h = ~0ull;
    u128 pr = ((u128)h << 64) | l;
    uint64_t got = red_v3((uint32_t)l, (uint32_t)(l >> 32), (uint32_t)h, (uint32_t)(h >> 32));
    if (got % P != (uint64_t)(pr % P)) { if (bad < 5) printf("BAD3 l=%016lx h=%016lx\n", l, h); bad++; }
    cnt++;
  }
  // gl::sqr3: the 128-bit square assembled from a0^2, a0 a1 (doubled with shifts) and a1^2 + addend + carry-in, word by word
  for (long it = 0; it < 40000000; it++) {
    uint64_t a = it < ne ? edge[it] : rnd();
    if (it % 7 == 0) a |= 0xFFFFFFFF00000000ull; if (it % 11 == 0) a |= 0xFFFFFFFFull; if (it % 13 == 0) a &= 0x80000000FFFFFFFFull;
    if (it % 17 == 0) a |= 0x8000000080000000ull;
    const uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32);
    const uint64_t pp = (uint64_t)a0 * a0, m = (uint64_t)a0 * a1;
    const uint32_t p0 = (uint32_t)pp, p1 = (uint32_t)(pp >> 32), m0 = (uint32_t)m, m1 = (uint32_t)(m >> 32);
    const uint32_t d0 = m0 << 1, d1 = (m1 << 1) | (m0 >> 31), d2 = m1 >> 31;
    const uint64_t s1 = (uint64_t)p1 + d0;                                              // add.cc
    const uint32_t w1 = (uint32_t)s1;
    const u128 q = (u128)((uint64_t)a1 * a1) + (((uint64_t)d2 << 32) | d1) + (s1 >> 32);   // madc.lo.cc / madc.hi
    if (q >> 64) { if (bad < 5) printf("BAD4 overflow a=%016lx\n", a); bad++; }
    const uint32_t w2 = (uint32_t)q, w3 = (uint32_t)(q >> 32);
    const u128 sq = (u128)a * a;
    if ((((u128)w3 << 96) | ((u128)w2 << 64) | ((u128)w1 << 32) | p0) != sq) { if (bad < 5) printf("BAD4 a=%016lx\n", a); bad++; }
    if (red_v3(p0, w1, w2, w3) % P != (uint64_t)(sq % P)) { if (bad < 5) printf("BAD5 a=%016lx\n", a); bad++; }
    cnt++;
  }
  printf("checked %ld, bad %ld\n", cnt, bad);
  return bad != 0;
}
