/*
 * pmt_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See pmt_oracle.h for scope and pinning.
 *
 * Every function cites the reference lines it restates.  "[UPSTREAM]" = plonky2 v0.1.3 @ 3b21b87d, the
 * un-vendored git dependency that implements PoseidonHash for the reference (Cargo.toml:7).
 */
#include "pmt_oracle.h"
#include "poseidon_constants.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
#define EPSILON 0xFFFFFFFFULL /* 2^32 - 1 = 2^64 mod p */

/* ------------------------------------------------------------------------------------------------------
 * Goldilocks field.  [UPSTREAM field/src/goldilocks_field.rs]: values are u64 and may be non-canonical;
 * to_canonical_u64 subtracts p once; reduce128 as published.
 * ---------------------------------------------------------------------------------------------------- */
uint64_t pmt_oracle_canonical(uint64_t x) { return x >= PMT_P ? x - PMT_P : x; }

static inline uint64_t add_no_canon(uint64_t x, uint64_t y) {
  uint64_t r = x + y;
  return r + EPSILON * (uint64_t)(r < x); /* cannot overflow twice when y is a reduce-intermediate */
}

/* [UPSTREAM field/src/goldilocks_field.rs reduce128], branch-free (upstream hints the borrow branch as unlikely and
 * uses an sbb trick on x86; a data-dependent branch here costs the scalar port 2x) */
static inline uint64_t reduce128(u128 x) {
  uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
  uint64_t hi_hi = hi >> 32, hi_lo = hi & EPSILON;
  uint64_t t0, r;
  uint64_t borrow = __builtin_sub_overflow(lo, hi_hi, &t0);
  t0 -= (0 - borrow) & EPSILON; /* borrow: 2^64 = eps (mod p) */
  uint64_t t1 = (hi_lo << 32) - hi_lo; /* hi_lo * eps */
  uint64_t carry = __builtin_add_overflow(t0, t1, &r);
  return r + ((0 - carry) & EPSILON); /* cannot overflow twice */
}
static inline uint64_t fadd(uint64_t a, uint64_t b) { /* a, b any u64 */
  return reduce128((u128)a + (u128)b);
}
static inline uint64_t fmul(uint64_t a, uint64_t b) { return reduce128((u128)a * (u128)b); }
uint64_t pmt_oracle_mul(uint64_t a, uint64_t b) { return pmt_oracle_canonical(fmul(a, b)); }

static inline uint64_t sbox7(uint64_t x) { /* [UPSTREAM hash/poseidon.rs sbox_monomial]: x^7 */
  uint64_t x2 = fmul(x, x), x4 = fmul(x2, x2), x3 = fmul(x, x2);
  return fmul(x3, x4);
}

/* [UPSTREAM hash/poseidon.rs mds_layer / mds_row_shf]:
 *   out[r] = sum_i state[(i + r) % 12] * MDS_MATRIX_CIRC[i] + state[r] * MDS_MATRIX_DIAG[r]            */
static void mds_layer(uint64_t s[12]) {
  uint64_t out[12];
  for (int r = 0; r < 12; r++) {
    u128 acc = 0;
    for (int i = 0; i < 12; i++) acc += (u128)s[(i + r) % 12] * PMT_MDS_CIRC[i];
    acc += (u128)s[r] * PMT_MDS_DIAG[r];
    out[r] = reduce128(acc);
  }
  memcpy(s, out, sizeof out);
}

/* [UPSTREAM hash/poseidon.rs Poseidon::poseidon] in its *specification* form: 4 full, 22 partial, 4 full rounds;
 * every round adds 12 constants, applies x^7 (all lanes / lane 0), then the MDS layer. */
void pmt_oracle_permute(uint64_t s[12]) {
  for (int r = 0; r < PMT_ROUNDS; r++) {
    for (int i = 0; i < 12; i++) s[i] = fadd(s[i], PMT_RC[12 * r + i]);
    if (r < PMT_FULL_HALF || r >= PMT_FULL_HALF + PMT_PARTIAL) {
      for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
    } else {
      s[0] = sbox7(s[0]);
    }
    mds_layer(s);
  }
  for (int i = 0; i < 12; i++) s[i] = pmt_oracle_canonical(s[i]);
}

/* ---- CPU-baseline speed path ------------------------------------------------------------------------------------
 * Same permutation, engineered the way upstream's scalar x86 path is ([UPSTREAM hash/poseidon.rs mds_layer via
 * u64 hi/lo accumulation, fast partial rounds]): the MDS layer works on 32-bit halves in i64 (sums < 2^45, one
 * reduction per lane) as a shift-and-add convolution, partial rounds use the sparse-matrix form.  Tables re-derived by
 * tools/gen_constants.py; equality with the naive form is a unit test.  Used for CPU-baseline timing. */
static inline uint64_t reduce96(uint64_t lo, uint64_t hi /* < 2^32 */) {
  return add_no_canon(lo, hi * EPSILON);
}

/* One 32-bit half of the state through the circulant part of the MDS matrix, as a length-12 cyclic convolution split by
 * t^12 - 1 = (t^3 - 1)(t^3 + 1)(t^6 + 1): the residues of the kernel are powers of two (64,128,64 | -4,-32,8 |
 * 4,-8,32,2,-2,-2; the inverse's /4 and /2 folded in), so the layer is shifts and adds on i64 -- the idea of upstream's
 * frequency-domain mds_multiply_freq.  Exact: every intermediate is an integer below 2^45. */
static inline void mds_half_freq(const int64_t x[12], int64_t y[12]) {
  int64_t e[6], d[6], a[3], b[3], A[3], B[3], D[6];
  for (int j = 0; j < 6; j++) { e[j] = x[j] + x[j + 6]; d[j] = x[j] - x[j + 6]; }
  for (int j = 0; j < 3; j++) { a[j] = e[j] + e[j + 3]; b[j] = e[j] - e[j + 3]; }
  /* cyclic 3x3 with (16, 32, 16) */
  A[0] = 16 * a[0] + 32 * a[2] + 16 * a[1];
  A[1] = 16 * a[1] + 32 * a[0] + 16 * a[2];
  A[2] = 16 * a[2] + 32 * a[1] + 16 * a[0];
  /* negacyclic 3x3 with (-1, -8, 2) */
  B[0] = -b[0] + 8 * b[2] - 2 * b[1];
  B[1] = -b[1] - 8 * b[0] - 2 * b[2];
  B[2] = -b[2] - 8 * b[1] + 2 * b[0];
  /* negacyclic 6x6 with (2, -4, 16, 1, -1, -1) */
  D[0] = 2 * d[0] + 4 * d[5] - 16 * d[4] - d[3] + d[2] + d[1];
  D[1] = 2 * d[1] - 4 * d[0] - 16 * d[5] - d[4] + d[3] + d[2];
  D[2] = 2 * d[2] - 4 * d[1] + 16 * d[0] - d[5] + d[4] + d[3];
  D[3] = 2 * d[3] - 4 * d[2] + 16 * d[1] + d[0] + d[5] + d[4];
  D[4] = 2 * d[4] - 4 * d[3] + 16 * d[2] + d[1] - d[0] + d[5];
  D[5] = 2 * d[5] - 4 * d[4] + 16 * d[3] + d[2] - d[1] - d[0];
  for (int j = 0; j < 3; j++) {
    int64_t p = A[j] + B[j], m = A[j] - B[j];
    y[j] = p + D[j]; y[j + 6] = p - D[j]; y[j + 3] = m + D[j + 3]; y[j + 9] = m - D[j + 3];
  }
}

static inline void mds_layer_fast(uint64_t s[12], const uint64_t *add) {
  int64_t lo[12], hi[12], L[12], H[12];
  for (int i = 0; i < 12; i++) { lo[i] = (int64_t)(s[i] & EPSILON); hi[i] = (int64_t)(s[i] >> 32); }
  mds_half_freq(lo, L);
  mds_half_freq(hi, H);
  L[0] += 8 * lo[0]; H[0] += 8 * hi[0];
  for (int r = 0; r < 12; r++) {
    /* value = L + 2^32 H, H = hh 2^32 + hl  ->  L + 2^32 hl + (2^32 - 1) hh */
    uint64_t l = (uint64_t)L[r], h = (uint64_t)H[r];
    u128 v = (u128)l + ((u128)(h & EPSILON) << 32) + (u128)(h >> 32) * EPSILON + (add ? add[r] : 0);
    s[r] = reduce96((uint64_t)v, (uint64_t)(v >> 64));
  }
}

/* x^7 on all 12 lanes, staged so that the 12 independent multiplications of each stage overlap in the pipeline */
static inline void sbox_layer_fast(uint64_t s[12]) {
  uint64_t x2[12], x3[12], x4[12];
  for (int i = 0; i < 12; i++) x2[i] = fmul(s[i], s[i]);
  for (int i = 0; i < 12; i++) x4[i] = fmul(x2[i], x2[i]);
  for (int i = 0; i < 12; i++) x3[i] = fmul(x2[i], s[i]);
  for (int i = 0; i < 12; i++) s[i] = fmul(x3[i], x4[i]);
}

static inline uint64_t dot12(const uint64_t *x, const uint64_t *k, int n) {
  u128 lo = 0, hi = 0; /* sum of n <= 12 128-bit products kept as two 68-bit halves */
  for (int i = 0; i < n; i++) {
    u128 pr = (u128)x[i] * k[i];
    lo += (uint64_t)pr;
    hi += (uint64_t)(pr >> 64);
  }
  return fadd(reduce128(lo), reduce128(hi * (u128)EPSILON));
}

void pmt_oracle_permute_fast(uint64_t s[12]) {
  for (int i = 0; i < 12; i++) s[i] = fadd(s[i], PMT_RC[i]);
  for (int r = 0; r < PMT_FULL_HALF; r++) {
    sbox_layer_fast(s);
    mds_layer_fast(s, r + 1 < PMT_FULL_HALF ? &PMT_RC[12 * (r + 1)] : PMT_FP_FIRST_RC);
  }
  {
    uint64_t t[12];
    t[0] = s[0];
    for (int a = 1; a < 12; a++) t[a] = dot12(s + 1, &PMT_FP_INIT[11 * (a - 1)], 11);
    memcpy(s, t, sizeof t);
  }
  for (int k = 0; k < PMT_PARTIAL; k++) {
    uint64_t x0 = fadd(sbox7(s[0]), PMT_FP_POST_RC[k]);
    uint64_t d = fadd(dot12(s + 1, &PMT_FP_W_HAT[11 * k], 11), reduce128((u128)x0 * PMT_FP_M00));
    for (int i = 1; i < 12; i++) s[i] = reduce128((u128)x0 * PMT_FP_V[11 * k + (i - 1)] + s[i]);
    s[0] = d;
  }
  for (int i = 0; i < 12; i++) s[i] = fadd(s[i], PMT_RC[12 * (PMT_FULL_HALF + PMT_PARTIAL) + i]);
  for (int r = PMT_FULL_HALF + PMT_PARTIAL; r < PMT_ROUNDS; r++) {
    sbox_layer_fast(s);
    mds_layer_fast(s, r + 1 < PMT_ROUNDS ? &PMT_RC[12 * (r + 1)] : NULL);
  }
  for (int i = 0; i < 12; i++) s[i] = pmt_oracle_canonical(s[i]);
}

typedef void (*perm_fn)(uint64_t *);

/* ------------------------------------------------------------------------------------------------------
 * Hasher.  [UPSTREAM hash/hashing.rs compress / hash_n_to_m_no_pad, plonk/config.rs Hasher::hash_or_noop]
 * ---------------------------------------------------------------------------------------------------- */
static void two_to_one_with(perm_fn perm, const uint64_t l[4], const uint64_t r[4], uint64_t out[4]) {
  uint64_t s[12] = {l[0], l[1], l[2], l[3], r[0], r[1], r[2], r[3], 0, 0, 0, 0};
  perm(s);
  memcpy(out, s, 32);
}
void pmt_oracle_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]) {
  two_to_one_with(pmt_oracle_permute, l, r, out);
}

static void hash_no_pad_with(perm_fn perm, const uint64_t *in, size_t n, uint64_t out[4]) {
  uint64_t s[12] = {0};
  for (size_t off = 0; off < n; off += 8) { /* overwrite mode, rate 8, no padding */
    size_t len = n - off < 8 ? n - off : 8;
    memcpy(s, in + off, len * 8);
    perm(s);
  }
  for (int i = 0; i < 4; i++) out[i] = pmt_oracle_canonical(s[i]);
}
void pmt_oracle_hash_no_pad(const uint64_t *in, size_t n, uint64_t out[4]) {
  hash_no_pad_with(pmt_oracle_permute, in, n, out);
}

static void hash_or_noop_with(perm_fn perm, const uint64_t *in, size_t n, uint64_t out[4]) {
  if (n <= 4) { /* inputs.len() * 8 <= 32 bytes: identity with zero padding */
    for (size_t i = 0; i < 4; i++) out[i] = i < n ? pmt_oracle_canonical(in[i]) : 0;
  } else {
    hash_no_pad_with(perm, in, n, out);
  }
}
void pmt_oracle_hash_or_noop(const uint64_t *in, size_t n, uint64_t out[4]) {
  hash_or_noop_with(pmt_oracle_permute, in, n, out);
}

static int log2_strict(size_t n) { /* plonky2_util::log2_strict: panics unless power of two */
  if (n == 0 || (n & (n - 1))) return -1;
  int l = 0;
  while (((size_t)1 << l) < n) l++;
  return l;
}

/* ------------------------------------------------------------------------------------------------------
 * simple_merkle_tree.rs
 * ---------------------------------------------------------------------------------------------------- */
/* MerkleTree::build, simple_merkle_tree.rs:28-51 (next_level_hashes :21-25).  n must be a power of two >= 2
 * (log2_strict :30 panics otherwise; n == 1 underflows at :38). */
int pmt_oracle_simple_tree_build(const uint64_t *leaves, size_t n, uint64_t *levels_out, uint64_t root_out[4]) {
  int count_levels = log2_strict(n);
  if (count_levels < 1) return -1;
  for (size_t i = 0; i < n; i++) pmt_oracle_hash_or_noop(leaves + i, 1, levels_out + 4 * i); /* :33 */
  uint64_t *cur = levels_out;
  size_t m = n;
  for (int i = 0; i < count_levels - 1; i++) { /* :38-41 */
    uint64_t *next = cur + 4 * m;
    for (size_t k = 0; k < m / 2; k++) pmt_oracle_two_to_one(cur + 8 * k, cur + 8 * k + 4, next + 4 * k); /* :23 */
    cur = next;
    m /= 2;
  }
  pmt_oracle_two_to_one(cur, cur + 4, root_out); /* :45 */
  return 0;
}

static const uint64_t *simple_level(const uint64_t *levels, size_t n, int level) {
  size_t off = 0, m = n;
  for (int i = 0; i < level; i++) { off += m; m /= 2; }
  return levels + 4 * off;
}

/* get_merkle_proof, simple_merkle_tree.rs:55-74 */
int pmt_oracle_simple_tree_proof(const uint64_t *levels, size_t n, size_t leaf_index, uint64_t *proof_out) {
  int count_levels = log2_strict(n);
  if (count_levels < 1 || leaf_index >= n) return -1; /* assert :56 */
  size_t idx = leaf_index;
  for (int i = 0; i < count_levels; i++) {
    const uint64_t *lvl = simple_level(levels, n, i);
    size_t sel = (idx & 1) ? idx - 1 : idx + 1; /* :64-68 */
    memcpy(proof_out + 4 * i, lvl + 4 * sel, 32);
    idx /= 2;
  }
  return 0;
}

/* get_in_between_hashes, simple_merkle_tree.rs:76-87 */
int pmt_oracle_simple_tree_in_between(const uint64_t *levels, const uint64_t root[4], size_t n, size_t leaf_index,
                                      uint64_t *out) {
  int count_levels = log2_strict(n);
  if (count_levels < 1 || leaf_index >= n) return -1; /* assert :77 */
  size_t idx = leaf_index / 2;
  int k = 0;
  for (int i = 1; i < count_levels; i++) {
    memcpy(out + 4 * k++, simple_level(levels, n, i) + 4 * idx, 32);
    idx /= 2;
  }
  memcpy(out + 4 * k, root, 32);
  return 0;
}

static int digest_eq(const uint64_t a[4], const uint64_t b[4]) {
  for (int i = 0; i < 4; i++)
    if (pmt_oracle_canonical(a[i]) != pmt_oracle_canonical(b[i])) return 0;
  return 1;
}

/* verify_merkle_proof, simple_merkle_tree.rs:91-109 */
int pmt_oracle_simple_tree_verify(uint64_t leaf, size_t leaf_index, const uint64_t root[4], const uint64_t *hashes,
                                  size_t n_hashes) {
  uint64_t cur[4], nxt[4];
  pmt_oracle_hash_or_noop(&leaf, 1, cur); /* :93 */
  size_t idx = leaf_index;
  for (size_t i = 0; i < n_hashes; i++) {
    if ((idx & 1) == 0) pmt_oracle_two_to_one(cur, hashes + 4 * i, nxt); /* :100 */
    else pmt_oracle_two_to_one(hashes + 4 * i, cur, nxt);                /* :102 */
    memcpy(cur, nxt, 32);
    idx /= 2;
  }
  return digest_eq(cur, root); /* :108 */
}

/* ------------------------------------------------------------------------------------------------------
 * merkle_mountain_ranges.rs
 * ---------------------------------------------------------------------------------------------------- */
/* get_heights_bitmap_for_mmr_size, merkle_mountain_ranges.rs:39-81 */
void pmt_oracle_mmr_heights_bitmap(size_t mmr_size, uint64_t *peaks_out, size_t *rem_out) {
  if (mmr_size == 0) { *peaks_out = 0; *rem_out = 0; return; }
  size_t subtree_size = (~(size_t)0) >> __builtin_clzll((unsigned long long)mmr_size); /* :44 */
  size_t updated = mmr_size;
  uint64_t peaks = 0;
  while (subtree_size > 0) { /* :68-78 */
    peaks <<= 1;
    if (updated >= subtree_size) { peaks |= 1; updated -= subtree_size; }
    subtree_size >>= 1;
  }
  *peaks_out = peaks;
  *rem_out = updated;
}

/* get_mmr_index, merkle_mountain_ranges.rs:257-270 (i32 arithmetic in the reference => index < 2^30) */
size_t pmt_oracle_mmr_index(size_t leaf_normal_index) {
  size_t index = leaf_normal_index, res = 0;
  unsigned height = 1;
  while (index > 0) {
    if (index & 1) res += ((size_t)1 << height) - 1;
    height++;
    index >>= 1;
  }
  return res;
}

/* MMR::add_leaf, merkle_mountain_ranges.rs:89-120 */
void pmt_oracle_mmr_add_leaf(uint64_t *elements, size_t *len, uint64_t leaf) {
  uint64_t next_hash[4];
  pmt_oracle_hash_or_noop(&leaf, 1, next_hash); /* :91 / :96 */
  if (*len == 0) { memcpy(elements, next_hash, 32); *len = 1; return; }
  uint64_t peaks; size_t rem;
  pmt_oracle_mmr_heights_bitmap(*len, &peaks, &rem); /* :102 */
  size_t current_pos = *len;
  memcpy(elements + 4 * (*len)++, next_hash, 32); /* :104 */
  unsigned height = 1;
  while (peaks > 0) { /* :106-119 */
    if (peaks & 1) {
      size_t prev_peak_index = current_pos - (((size_t)1 << height) - 1);
      uint64_t tmp[4];
      pmt_oracle_two_to_one(elements + 4 * prev_peak_index, next_hash, tmp); /* :111 */
      memcpy(next_hash, tmp, 32);
      memcpy(elements + 4 * (*len)++, next_hash, 32);
    } else {
      break;
    }
    peaks >>= 1;
    height++;
    current_pos++;
  }
}

/* MMR::get_peaks, merkle_mountain_ranges.rs:179-200 (u32 arithmetic at :184 => len < 2^32) */
size_t pmt_oracle_mmr_peaks(const uint64_t *elements, size_t len, uint64_t *peaks_out) {
  if (len == 0) return 0;
  size_t max_tree_size = 0xFFFFFFFFu >> __builtin_clz((unsigned)len);
  size_t current_index = len, peak_pos = 0, k = 0;
  while (max_tree_size > 0) {
    if (current_index >= max_tree_size) {
      peak_pos += max_tree_size;
      memcpy(peaks_out + 4 * k++, elements + 4 * (peak_pos - 1), 32);
      current_index -= max_tree_size;
    }
    max_tree_size >>= 1;
  }
  return k;
}

/* MMR::bagging_the_peaks, merkle_mountain_ranges.rs:122-127 */
void pmt_oracle_mmr_bag(const uint64_t *elements, size_t len, uint64_t root_out[4]) {
  uint64_t peaks[4 * 64];
  size_t k = pmt_oracle_mmr_peaks(elements, len, peaks);
  pmt_oracle_hash_or_noop(peaks, 4 * k, root_out); /* :125 */
}

/* MMR::get_subtree_proof_elm + add_right_elm, merkle_mountain_ranges.rs:129-176 */
size_t pmt_oracle_mmr_subtree_proof(const uint64_t *elements, size_t len, size_t mmr_index, uint64_t *siblings_out,
                                    uint8_t *on_left_out) {
  size_t n = 0, curr_index = mmr_index;
  int intree = 1;
  unsigned height = 0;
  while (intree) {
    size_t span = ((size_t)1 << (height + 1)) - 1;
    int took_left = 0;
    if (curr_index >= span) { /* :158 */
      size_t prev_elm_index = curr_index - span;
      uint64_t pk; size_t rem;
      pmt_oracle_mmr_heights_bitmap(prev_elm_index, &pk, &rem);
      if (rem == height) { /* :161 */
        memcpy(siblings_out + 4 * n, elements + 4 * prev_elm_index, 32);
        on_left_out[n++] = 1;
        curr_index += 1;
        took_left = 1;
      }
    }
    if (!took_left) { /* add_right_elm :129-144 */
      size_t next_elm_index = curr_index + span;
      if (len >= 1 && next_elm_index < len - 1) {
        memcpy(siblings_out + 4 * n, elements + 4 * next_elm_index, 32);
        on_left_out[n++] = 0;
        curr_index = next_elm_index + 1;
      } else {
        intree = 0;
      }
    }
    height++;
  }
  return n;
}

/* MMR_proof::verify, merkle_mountain_ranges.rs:232-252 */
int pmt_oracle_mmr_verify(uint64_t leaf, const uint64_t root[4], const uint64_t *siblings, const uint8_t *on_left,
                          size_t path_len, const uint64_t *peaks, size_t n_peaks) {
  uint64_t cur[4], nxt[4];
  pmt_oracle_hash_or_noop(&leaf, 1, cur); /* :233 */
  for (size_t i = 0; i < path_len; i++) {
    if (on_left[i]) pmt_oracle_two_to_one(siblings + 4 * i, cur, nxt); /* :238 */
    else pmt_oracle_two_to_one(cur, siblings + 4 * i, nxt);            /* :240 */
    memcpy(cur, nxt, 32);
  }
  int found = 0;
  for (size_t k = 0; k < n_peaks; k++) found |= digest_eq(peaks + 4 * k, cur);
  if (!found) return -1; /* assert! :245 */
  uint64_t calc[4];
  pmt_oracle_hash_or_noop(peaks, 4 * n_peaks, calc); /* :249 */
  return digest_eq(calc, root);
}

/* ------------------------------------------------------------------------------------------------------
 * plonky2 MerkleTree::new.  [UPSTREAM hash/merkle_tree.rs fill_subtree / fill_digests_buf / MerkleTree::new]
 * Layout of a subtree's digest buffer: left recursive output || left child digest || right child digest ||
 * right recursive output; the subtree's own root is returned to the caller (parent or cap).
 * ---------------------------------------------------------------------------------------------------- */
#define PAR_CUTOFF 512 /* leaves below which the fork-join stops spawning tasks */

static void fill_subtree(perm_fn perm, uint64_t *digests, size_t n_digests, const uint64_t *leaves, size_t n_leaves,
                         size_t w, uint64_t out[4], int parallel) {
  if (n_digests == 0) {
    hash_or_noop_with(perm, leaves, w, out);
    return;
  }
  size_t half = n_digests / 2;
  uint64_t *left_buf = digests, *left_digest = digests + 4 * (half - 1);
  uint64_t *right_digest = digests + 4 * half, *right_buf = digests + 4 * (half + 1);
  uint64_t l[4], r[4];
  if (parallel && n_leaves > PAR_CUTOFF) {
#pragma omp task shared(l) firstprivate(perm, left_buf, half, leaves, n_leaves, w)
    fill_subtree(perm, left_buf, half - 1, leaves, n_leaves / 2, w, l, 1);
#pragma omp task shared(r) firstprivate(perm, right_buf, half, leaves, n_leaves, w)
    fill_subtree(perm, right_buf, half - 1, leaves + (n_leaves / 2) * w, n_leaves / 2, w, r, 1);
#pragma omp taskwait
  } else {
    fill_subtree(perm, left_buf, half - 1, leaves, n_leaves / 2, w, l, 0);
    fill_subtree(perm, right_buf, half - 1, leaves + (n_leaves / 2) * w, n_leaves / 2, w, r, 0);
  }
  memcpy(left_digest, l, 32);
  memcpy(right_digest, r, 32);
  two_to_one_with(perm, l, r, out);
}

int pmt_oracle_merkle_tree_new(const uint64_t *leaves, size_t n, size_t w, unsigned cap_height, uint64_t *digests_out,
                               uint64_t *cap_out, int threads, int use_fast) {
  int lg = log2_strict(n);
  if (lg < 0 || (int)cap_height > lg) return -1; /* assert in MerkleTree::new */
  perm_fn perm = use_fast ? pmt_oracle_permute_fast : pmt_oracle_permute;
  size_t n_cap = (size_t)1 << cap_height;
  size_t num_digests = 2 * (n - n_cap);
  size_t sub_digests = num_digests >> cap_height, sub_leaves = n >> cap_height;
  int parallel = threads > 1;
#ifndef _OPENMP
  parallel = 0;
#endif
  if (!parallel) {
    for (size_t c = 0; c < n_cap; c++)
      fill_subtree(perm, digests_out + 4 * c * sub_digests, sub_digests, leaves + c * sub_leaves * w, sub_leaves, w,
                   cap_out + 4 * c, 0);
    return 0;
  }
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#pragma omp single
  {
    for (size_t c = 0; c < n_cap; c++) {
#pragma omp task firstprivate(c)
      fill_subtree(perm, digests_out + 4 * c * sub_digests, sub_digests, leaves + c * sub_leaves * w, sub_leaves, w,
                   cap_out + 4 * c, 1);
    }
#pragma omp taskwait
  }
#endif
  return 0;
}

/* [UPSTREAM hash/merkle_tree.rs MerkleTree::prove] */
int pmt_oracle_merkle_prove(const uint64_t *digests, size_t n, unsigned cap_height, size_t leaf_index,
                            uint64_t *siblings_out) {
  int lg = log2_strict(n);
  if (lg < 0 || (int)cap_height > lg || leaf_index >= n) return -1;
  unsigned num_layers = (unsigned)lg - cap_height;
  size_t tree_len = (2 * (n - ((size_t)1 << cap_height))) >> cap_height;
  const uint64_t *digest_tree = digests + 4 * tree_len * (leaf_index >> num_layers);
  size_t pair_index = leaf_index & (((size_t)1 << num_layers) - 1);
  for (unsigned i = 0; i < num_layers; i++) {
    size_t parity = pair_index & 1;
    pair_index >>= 1;
    size_t siblings_index = (pair_index << (i + 1)) + ((size_t)1 << i) - 1;
    size_t sibling_index = 2 * siblings_index + (1 - parity);
    memcpy(siblings_out + 4 * i, digest_tree + 4 * sibling_index, 32);
  }
  return 0;
}

/* [UPSTREAM hash/merkle_proofs.rs verify_merkle_proof_to_cap] */
int pmt_oracle_merkle_verify_to_cap(const uint64_t *leaf, size_t w, size_t leaf_index, const uint64_t *cap,
                                    unsigned cap_height, const uint64_t *siblings, size_t n_siblings) {
  uint64_t cur[4], nxt[4];
  size_t index = leaf_index;
  pmt_oracle_hash_or_noop(leaf, w, cur);
  for (size_t i = 0; i < n_siblings; i++) {
    if (index & 1) pmt_oracle_two_to_one(siblings + 4 * i, cur, nxt);
    else pmt_oracle_two_to_one(cur, siblings + 4 * i, nxt);
    memcpy(cur, nxt, 32);
    index >>= 1;
  }
  if (index >= ((size_t)1 << cap_height)) return 0;
  return digest_eq(cur, cap + 4 * index);
}

int pmt_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
