// poseidon_freq.cuh -- the Poseidon MDS layer as a length-12 cyclic convolution in the "frequency" domain, on the fp64 pipe.
//
// Same arithmetic object as poseidon.cuh mds_layer_dfma2 / permute_paired ([UPSTREAM plonky2 hash/poseidon.rs mds_layer,
// hash/poseidon_goldilocks.rs MDS_MATRIX_CIRC / MDS_MATRIX_DIAG], under PoseidonHash at
// /root/reference/src/simple_merkle_tree/simple_merkle_tree.rs:23,33 and /root/reference/src/mmr/merkle_mountain_ranges.rs:96,111),
// fewer instructions: the MDS matrix is M = C + 8 e0 e0^T with C circulant, and
//     t^12 - 1 = (t^3 - 1)(t^3 + 1)(t^6 + 1)
// splits y = C x with additions only into a cyclic 3 x 3, a negacyclic 3 x 3 and a negacyclic 6 x 6 product:
// 18 + 54 + 18 = 90 DADD/DFMA per 32-bit half of the state instead of 144 DFMA.  The inverse transform's halvings are
// folded into the product coefficients (multiples of 1/4: exact), and 2^52 + the round constants' halves are added last
// so the mantissa of each output is the integer.
// DFMA holds the issue port of a B200 SM sub-partition for 2.2 cycles and overlaps poorly with integer work
// (profiles/pipes_r1.jsonl), so fp64 instructions saved are cycles saved.
//
// Paired partial rounds (poseidon.cuh permute_paired: z = A s' + M[:,0] x + K) become
//     z = C^2 s' + C[:,0] (8 s'_0 + x - y0) + 8 x e0 + (C c1 + c2),     y0 = M[0,:] s' + c1_0,   x = sbox(y0)
// with C^2 circulant as well: 260 fp64 operations per pair instead of 336.
//
// tools/gen_freq_constants.py derives the tables, checks both identities against the matrix forms in exact rational
// arithmetic and PROVES that every intermediate value below (same operation order) is exactly representable for all
// inputs; tests/cpp/check_freq.cpp runs this very header on the host against the specification-form CPU permutation.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define PMT_FQ_HD __device__ __forceinline__
#else
#include <cmath>
#define PMT_FQ_HD static inline
#define PMT_FQ_QUAL static const double
#endif
#include "poseidon_freq_constants.cuh"

namespace poseidon {
namespace freq {

static constexpr int W = 12;
static constexpr double MAGIC = 4503599627370496.0;  // 2^52

// x -> [a 3 | b 3 | v 6]: residues mod t^3 - 1, t^3 + 1, t^6 + 1
PMT_FQ_HD void fwd(double (&x)[W]) {
  double u[6], v[6];
#pragma unroll
  for (int j = 0; j < 6; j++) { u[j] = x[j] + x[j + 6]; v[j] = x[j] - x[j + 6]; }
#pragma unroll
  for (int j = 0; j < 3; j++) { x[j] = u[j] + u[j + 3]; x[3 + j] = u[j] - u[j + 3]; }
#pragma unroll
  for (int j = 0; j < 6; j++) x[6 + j] = v[j];
}

// the three residue products; pc = [ka 3 | kb 3 | kv 6]: the kernel's residues, pre-scaled by 1/4, 1/4, 1/2.  The product
// matrices are Toeplitz in them (cyclic, negacyclic, negacyclic); the signs are operand negations (free), so a layer
// needs only 12 distinct constants -- they live in uniform registers (DFMA has no constant-bank operand form; a table
// of all 54 signed entries made ptxas spill uniform registers into 100 vector registers).
PMT_FQ_HD void mul(const double (&f)[W], double (&o)[W], const double* __restrict__ pc) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double acc = f[0] * pc[k];
#pragma unroll
    for (int i = 1; i < 3; i++) acc = fma(f[i], pc[(k - i + 3) % 3], acc);
    o[k] = acc;
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double acc = f[3] * pc[3 + k];
#pragma unroll
    for (int i = 1; i < 3; i++) acc = i <= k ? fma(f[3 + i], pc[3 + k - i], acc) : fma(-f[3 + i], pc[3 + k - i + 3], acc);
    o[3 + k] = acc;
  }
#pragma unroll
  for (int k = 0; k < 6; k++) {
    double acc = f[6] * pc[6 + k];
#pragma unroll
    for (int i = 1; i < 6; i++) acc = i <= k ? fma(f[6 + i], pc[6 + k - i], acc) : fma(-f[6 + i], pc[6 + k - i + 6], acc);
    o[6 + k] = acc;
  }
}

// [a | b | v] (pre-scaled) -> the 12 outputs
PMT_FQ_HD void inv(const double (&o)[W], double (&y)[W]) {
  double uu[6];
#pragma unroll
  for (int j = 0; j < 3; j++) { uu[j] = o[j] + o[3 + j]; uu[3 + j] = o[j] - o[3 + j]; }
#pragma unroll
  for (int j = 0; j < 6; j++) { y[j] = uu[j] + o[6 + j]; y[j + 6] = uu[j] - o[6 + j]; }
}

// One 32-bit half of a full MDS layer.  x: the 12 input halves (destroyed); km[STRIDE * r] = 2^52 + this half of the
// constant the layer adds to lane r.  out[r] = 2^52 + (M x + constants)_r.
template <int STRIDE>
PMT_FQ_HD void full_layer_half(double (&x)[W], const double* __restrict__ km, double (&out)[W]) {
  const double x0 = x[0];
  double o[W];
  fwd(x);
  mul(x, o, PMT_FQ_P1);
  inv(o, out);
  out[0] += fma(x0, 8.0, km[0]);   // the diagonal term 8 x_0 rides on lane 0's constant
#pragma unroll
  for (int j = 1; j < W; j++) out[j] += km[STRIDE * j];
}

// First part of one half of a PAIR of partial rounds, everything that does not need x = sbox(y0):
//   l0m = 2^52 + this half of y0 = M[0,:] s' + c1_0    (c1m = 2^52 + the half of c1_0)
//   y   = (C^2 s') half
// x: the 12 halves of s' (destroyed).  x0_out keeps s'_0's half for the second part.
PMT_FQ_HD void pair_half_begin(double (&x)[W], double c1m, double& l0m, double (&y)[W], double& x0_out) {
  double acc = c1m;
#pragma unroll
  for (int i = 0; i < W; i++) acc = fma(x[i], PMT_FQ_CIRC13[i == 0 ? 12 : i], acc);
  l0m = acc;
  x0_out = x[0];
  double o[W];
  fwd(x);
  mul(x, o, PMT_FQ_P2);
  inv(o, y);
}

// Second part: xs = this half of x = sbox(y0).  y (from pair_half_begin) becomes 2^52 + z's half.
// kpm[STRIDE * j] = 2^52 + this half of (C c1 + c2)_j.
template <int STRIDE>
PMT_FQ_HD void pair_half_end(double (&y)[W], double xs, double x0, double l0m, const double* __restrict__ kpm) {
  const double d = fma(x0, 8.0, xs) - (l0m - MAGIC);   // 8 s'_0 + x - y0: an exact integer, |d| < 2^42
#pragma unroll
  for (int j = 0; j < W; j++) {
    double t = fma(d, PMT_FQ_CIRC13[(W - j) % W], kpm[STRIDE * j]);   // C[j,0]
    if (j == 0) t = fma(xs, 8.0, t);
    y[j] += t;
  }
}

}  // namespace freq
}  // namespace poseidon
