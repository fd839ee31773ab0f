// host check of poseidon::combine_magic_c (plonky2_merkle_trees_b200/csrc/poseidon.cuh) against exact arithmetic:
// gcc -O2 -o check_combine tools/check_combine.c && ./check_combine
#include <stdint.h>
#include <stdio.h>
typedef unsigned __int128 u128;
static const uint64_t P = 0xFFFFFFFF00000001ull;
static uint64_t combine_c(uint64_t Lb, uint64_t Hb) {
  const uint32_t l0 = (uint32_t)Lb, L1 = (uint32_t)(Lb >> 32), h0 = (uint32_t)Hb, H1 = (uint32_t)(Hb >> 32);
  const int64_t lo = (int64_t)(uint64_t)l0 - (int64_t)(uint64_t)H1 + 0x43300000ll;
  const int64_t hi = (int64_t)(uint64_t)L1 + (int64_t)(uint64_t)h0 + (int64_t)(uint64_t)H1 + (lo >> 32) - 0x86600000ll;
  const int64_t n = hi >> 32;
  return (((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo) + ((uint64_t)n << 32) - (uint64_t)n;
}
static uint64_t rnd() { static uint64_t s = 88172645463325252ull; s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
int main() {
  const uint64_t M = (1ull << 52) - 1, MAGIC = 0x4330000000000000ull;
  uint64_t edge[] = {0, 1, 0xFFFFFFFFull, 0x100000000ull, 0xFFFFFFFFFull, M, M - 1, 0xFFFFF00000000ull, 0xFFFFFFFFFFFFull, 0x0000FFFF00000000ull, 0xFFFFEFFFFFFFFull};
  int ne = sizeof edge / sizeof *edge; long bad = 0, cnt = 0;
  for (long it = 0; it < 60000000; it++) {
    uint64_t l, h;
    if (it < ne * ne) { l = edge[it / ne]; h = edge[it % ne]; }
    else { l = rnd() & M; h = rnd() & M; if (it % 3 == 0) h |= 0xFFFFF00000000ull; if (it % 5 == 0) l &= 0xFFFFFFFFull; if (it % 7 == 0) l = M - (l & 0xFFFF); if (it % 11 == 0) h = (h & 0xFFFFF00000000ull) | (0xFFFFFFFFull - (h & 0xFF)); }
    uint64_t got = combine_c(MAGIC | l, MAGIC | h);
    uint64_t want = (uint64_t)(((u128)l + ((u128)h << 32)) % P);
    if (got % P != want) { if (bad < 5) printf("BAD l=%013lx h=%013lx got=%016lx want=%016lx\n", l, h, got, want); bad++; }
    cnt++;
  }
  printf("checked %ld, bad %ld\n", cnt, bad);
  return bad != 0;
}
