// check_freq.cpp -- host-side proof-by-execution for plonky2_merkle_trees_b200/csrc/poseidon_freq.cuh (tuning/verification
// aid, not product code).  The header is compiled here with g++ exactly as nvcc compiles it for the device (IEEE doubles,
// fma with one rounding), and checked
//   1. layer by layer against the integer matrix forms at EVERY corner of the input cube {0, 2^32 - 1}^12 (the layers are
//      linear, so each intermediate value takes its extreme magnitude at a corner) for all 16 + 22 constant sets;
//   2. as a whole permutation (full rounds + paired partial rounds, the device code's structure) against the oracle's
//      specification-form permutation on random and edge states.
// build: g++ -O2 -std=c++17 -ffp-contract=off -o /tmp/check_freq tests/cpp/check_freq.cpp -Loracle -lpmt_oracle -Wl,-rpath,$PWD/oracle
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../oracle/pmt_oracle.h"
#include "../../oracle/poseidon_constants.h"
#include "../../plonky2_merkle_trees_b200/csrc/poseidon_freq.cuh"

typedef unsigned __int128 u128;
using namespace poseidon::freq;
static const uint64_t P = 0xFFFFFFFF00000001ull;
static const uint64_t CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
static uint64_t M[12][12];

static uint64_t mulp(uint64_t a, uint64_t b) { return (uint64_t)((u128)a * b % P); }
static uint64_t sbox(uint64_t x) { x %= P; uint64_t x2 = mulp(x, x), x4 = mulp(x2, x2), x3 = mulp(x, x2); return mulp(x3, x4); }
static uint64_t from_magic(double v) {   // the integer in the mantissa of 2^52 + integer
  uint64_t b; memcpy(&b, &v, 8);
  if ((b >> 52) != 0x433) { printf("value %a left [2^52, 2^53)\n", v); exit(1); }
  return b & 0x000FFFFFFFFFFFFFull;
}
static uint64_t combine(double lo, double hi) { return (uint64_t)(((u128)from_magic(lo) + ((u128)from_magic(hi) << 32)) % P); }
static uint64_t rng_state = 0x1234567;
static uint64_t rnd() { uint64_t z = (rng_state += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }

static void permute_freq(uint64_t s[12]) {
  for (int i = 0; i < 12; i++) s[i] = (uint64_t)(((u128)s[i] + PMT_RC[i]) % P);
  auto full = [&](int layer) {
    double lo[12], hi[12], olo[12], ohi[12];
    for (int i = 0; i < 12; i++) { const uint64_t x = sbox(s[i]); lo[i] = (double)(uint32_t)x; hi[i] = (double)(uint32_t)(x >> 32); }
    double klo[12], khi[12];
    const int r = layer < 4 ? layer : 22 + layer;
    for (int i = 0; i < 12; i++) { const uint64_t c = r + 1 < 30 ? PMT_RC[12 * (r + 1) + i] : 0; klo[i] = MAGIC + (double)(uint32_t)c; khi[i] = MAGIC + (double)(uint32_t)(c >> 32); }
    full_layer_half<1>(lo, klo, olo);
    full_layer_half<1>(hi, khi, ohi);
    for (int i = 0; i < 12; i++) s[i] = combine(olo[i], ohi[i]);
  };
  for (int q = 0; q < 4; q++) full(q);
  for (int pair = 0; pair < 11; pair++) {
    const int r = 4 + 2 * pair;
    s[0] = sbox(s[0]);
    double lo[12], hi[12], ylo[12], yhi[12], l0lo, l0hi, x0lo, x0hi;
    for (int i = 0; i < 12; i++) { lo[i] = (double)(uint32_t)s[i]; hi[i] = (double)(uint32_t)(s[i] >> 32); }
    const uint64_t c1 = PMT_RC[12 * (r + 1)];
    pair_half_begin(lo, MAGIC + (double)(uint32_t)c1, l0lo, ylo, x0lo);
    pair_half_begin(hi, MAGIC + (double)(uint32_t)(c1 >> 32), l0hi, yhi, x0hi);
    const uint64_t x = sbox(combine(l0lo, l0hi));
    pair_half_end<2>(ylo, (double)(uint32_t)x, x0lo, l0lo, &PMT_FQ_KPAIR_DM[24 * pair]);
    pair_half_end<2>(yhi, (double)(uint32_t)(x >> 32), x0hi, l0hi, &PMT_FQ_KPAIR_DM[24 * pair + 1]);
    for (int i = 0; i < 12; i++) s[i] = combine(ylo[i], yhi[i]);
  }
  for (int q = 0; q < 4; q++) full(4 + q);
  for (int i = 0; i < 12; i++) s[i] %= P;
}

int main() {
  for (int r = 0; r < 12; r++) for (int c = 0; c < 12; c++) M[r][c] = CIRC[(c - r + 12) % 12] + (r == 0 && c == 0 ? 8 : 0);
  const uint64_t MAXH = 0xFFFFFFFFull;
  long checked = 0;
  // 1a. full layers at every corner
  for (int layer = 0; layer < 8; layer++) {
    const int r = layer < 4 ? layer : 22 + layer;
    for (int h = 0; h < 2; h++)
      for (unsigned mask = 0; mask < 4096; mask++) {
        double x[12], out[12]; uint64_t xi[12];
        for (int i = 0; i < 12; i++) { xi[i] = (mask >> i & 1) ? MAXH : 0; x[i] = (double)xi[i]; }
        double km[12];
        for (int q = 0; q < 12; q++) { const uint64_t c = r + 1 < 30 ? PMT_RC[12 * (r + 1) + q] : 0; km[q] = MAGIC + (double)(h ? c >> 32 : c & MAXH); }
        full_layer_half<1>(x, km, out);
        for (int q = 0; q < 12; q++) {
          const uint64_t c = r + 1 < 30 ? PMT_RC[12 * (r + 1) + q] : 0;
          uint64_t want = h ? c >> 32 : c & MAXH;
          for (int i = 0; i < 12; i++) want += M[q][i] * xi[i];
          if (from_magic(out[q]) != want) { printf("full layer %d half %d mask %u lane %d: %llu != %llu\n", layer, h, mask, q, (unsigned long long)from_magic(out[q]), (unsigned long long)want); return 1; }
          checked++;
        }
      }
  }
  // 1b. pairs at every corner, x half in {0, max, random}
  for (int pair = 0; pair < 11; pair++) {
    const int r = 4 + 2 * pair;
    for (int h = 0; h < 2; h++)
      for (unsigned mask = 0; mask < 4096; mask++)
        for (int xv = 0; xv < 3; xv++) {
          double x[12], y[12], l0m, x0; uint64_t xi[12];
          for (int i = 0; i < 12; i++) { xi[i] = (mask >> i & 1) ? MAXH : 0; x[i] = (double)xi[i]; }
          const uint64_t xs = xv == 0 ? 0 : xv == 1 ? MAXH : (rnd() & MAXH);
          const uint64_t c1 = PMT_RC[12 * (r + 1)];
          const uint64_t c1h = h ? c1 >> 32 : c1 & MAXH;
          pair_half_begin(x, MAGIC + (double)c1h, l0m, y, x0);
          uint64_t y0 = c1h;
          for (int i = 0; i < 12; i++) y0 += M[0][i] * xi[i];
          if (from_magic(l0m) != y0) { printf("pair %d y0 mismatch\n", pair); return 1; }
          pair_half_end<2>(y, (double)xs, x0, l0m, &PMT_FQ_KPAIR_DM[24 * pair + h]);
          // two rounds of the specification on this half: w = M s' + c1, w_0 := xs, z = M w + c2 (plain integers)
          uint64_t w[12];
          for (int q = 0; q < 12; q++) {
            const uint64_t c = PMT_RC[12 * (r + 1) + q];
            w[q] = h ? c >> 32 : c & MAXH;
            for (int i = 0; i < 12; i++) w[q] += M[q][i] * xi[i];
          }
          w[0] = xs;
          for (int q = 0; q < 12; q++) {
            const uint64_t c = PMT_RC[12 * (r + 2) + q];
            uint64_t want = h ? c >> 32 : c & MAXH;
            for (int i = 0; i < 12; i++) want += M[q][i] * w[i];
            if (from_magic(y[q]) != want) { printf("pair %d half %d mask %u lane %d: %llu != %llu\n", pair, h, mask, q, (unsigned long long)from_magic(y[q]), (unsigned long long)want); return 1; }
            checked++;
          }
        }
  }
  printf("layers: %ld outputs exact at all corners\n", checked);
  // 2. whole permutations against the oracle
  const uint64_t edge[] = {0, 1, 0xFFFFFFFFull, 0x100000000ull, P - 1, P, P + 1, ~0ull};
  long perms = 0;
  for (int t = 0; t < 200000; t++) {
    uint64_t a[12], b[12];
    for (int i = 0; i < 12; i++) a[i] = t < 64 ? edge[(t + i * (t / 8 + 1)) % 8] : rnd();
    memcpy(b, a, sizeof a);
    permute_freq(a);
    pmt_oracle_permute(b);
    for (int i = 0; i < 12; i++) if (a[i] != pmt_oracle_canonical(b[i])) { printf("permutation %d lane %d mismatch\n", t, i); return 1; }
    perms++;
  }
  printf("permutations: %ld equal to the oracle's specification form\nOK\n", perms);
  return 0;
}
