// ab_level.cu -- tuning harness (not part of the product library): times one big level of two_to_one (2^22 nodes, inputs
// larger than L2) for whatever permutation variant the -D flags select, and prints a checksum of the produced digests so
// that variants can be compared bit for bit.  PMT_PERM = 5 (default): the PRODUCT permutation (csrc/poseidon.cuh) through the
// product kernel k_level; 0..4: the round-1 forms kept in tools/experimental/.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 [-DPMT_...] -o ab_level tools/ab_level.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../plonky2_merkle_trees_b200/csrc/merkle_kernels.cuh"
#include "experimental/poseidon_quad.cuh"

#ifndef PMT_PERM
#define PMT_PERM 5
#endif

namespace ab {
using pmt::WIDTH; using pmt::Digest; using pmt::load_digest; using pmt::store_digest; using pmt::BLOCK;
// the production permutation (see DESIGN.md "Permutation variants" for the measurements behind this choice).
// PMT_PERM selects the form for A/B runs (tools/ab_level.cu): 0 = permute_fast (sparse partial rounds), 1 = permute_fused,
// 2 = permute_rounds (30 x S-box + DFMA MDS), 3 = permute_paired (partial rounds in pairs; round-1 production until the
// frequency form), 4 = permute_paired_freq (3 with the MDS layers as frequency-domain convolutions, poseidon_freq.cuh;
// production: 1.55 against 1.29 G permutations/s, profiles/ab_freq_r1.jsonl).
#ifndef PMT_SBOX_FMA_MASK
#define PMT_SBOX_FMA_MASK 0
#endif
#ifndef PMT_PART_FMA_MASK
#define PMT_PART_FMA_MASK 0
#endif
#ifndef PMT_MULADD_ALU
#define PMT_MULADD_ALU 1
#endif
#ifndef PMT_DOT_ALU
#define PMT_DOT_ALU 0
#endif
#ifndef PMT_CVT_I2F
#define PMT_CVT_I2F 1   // I2F.F64.U32 (conversion pipe) instead of the 2^52 magic-number subtraction (fma pipe)
#endif
#ifndef PMT_COMBINE_ALU
#define PMT_COMBINE_ALU 1
#endif
#ifndef PMT_COLUMN
#define PMT_COLUMN 0
#endif
#ifndef PMT_PIPE
#define PMT_PIPE 0
#endif
#ifndef PMT_PPIPE
#define PMT_PPIPE 0
#endif
#ifndef PMT_FQ_COMBINE
#define PMT_FQ_COMBINE 0   // recombination of the fp64 sums: 0 = ALU only, 1 = IMAD.WIDE folds, 2 = IMAD.WIDE in full layers only, 3 = in pairs only
#endif
#ifndef PMT_FQ_SPLIT
#define PMT_FQ_SPLIT 0   // 1: fence the high halves behind the low halves (measured slower: 1.51 against 1.55)
#endif
#ifndef PMT_COMPRESS_SPECIALISED
#define PMT_COMPRESS_SPECIALISED 0
#endif
template <bool CAP_ZERO, bool OUT4>
__device__ __forceinline__ void permute_impl(uint64_t (&s)[WIDTH]) {
#if PMT_PERM == 5
  poseidon::permute(s);
#elif PMT_PERM == 0
  poseidonx::permute_fast<true, true, 2, CAP_ZERO, OUT4>(s);
#elif PMT_PERM == 3
  poseidonx::permute_paired<PMT_SBOX_FMA_MASK, PMT_PART_FMA_MASK, PMT_COLUMN != 0, PMT_CVT_I2F != 0, PMT_COMBINE_ALU != 0, CAP_ZERO,
                           OUT4>(s);
#elif PMT_PERM == 4
  poseidonx::permute_paired_freq<PMT_SBOX_FMA_MASK, PMT_PART_FMA_MASK, PMT_CVT_I2F != 0, CAP_ZERO, OUT4, PMT_FQ_SPLIT, PMT_FQ_COMBINE>(s);
#elif PMT_PERM == 2
  poseidonx::permute_rounds<PMT_SBOX_FMA_MASK, PMT_PART_FMA_MASK, PMT_COLUMN != 0, PMT_CVT_I2F != 0, PMT_COMBINE_ALU != 0, CAP_ZERO,
                           OUT4>(s);
#else
  poseidonx::permute_fused<PMT_SBOX_FMA_MASK, PMT_PART_FMA_MASK, PMT_MULADD_ALU != 0, PMT_DOT_ALU != 0, PMT_CVT_I2F != 0,
                          PMT_COMBINE_ALU != 0, CAP_ZERO, OUT4, PMT_PIPE, PMT_PPIPE>(s);
#endif
}
__device__ __forceinline__ void permute(uint64_t (&s)[WIDTH]) { permute_impl<false, false>(s); }
// two_to_one: zero capacity lanes on entry, only the digest lanes are read afterwards
__device__ __forceinline__ void permute_compress(uint64_t (&s)[WIDTH]) {
  permute_impl<PMT_COMPRESS_SPECIALISED != 0, PMT_COMPRESS_SPECIALISED != 0>(s);
}


__device__ __forceinline__ Digest two_to_one_ab(const Digest& l, const Digest& r) {
  uint64_t s[WIDTH] = {l.v[0], l.v[1], l.v[2], l.v[3], r.v[0], r.v[1], r.v[2], r.v[3], 0, 0, 0, 0};
  permute_compress(s);
  Digest d;
#pragma unroll
  for (int i = 0; i < 4; i++) d.v[i] = gl::canonical(s[i]);
  return d;
}
#ifndef PMT_QUAD
#define PMT_QUAD 0
#endif
#ifndef PMT_QUAD_SBOX_FMA_MASK
#define PMT_QUAD_SBOX_FMA_MASK 0
#endif
#ifndef PMT_QUAD_PART_FMA_MASK
#define PMT_QUAD_PART_FMA_MASK 0
#endif
#ifndef PMT_QUAD_COMBINE_ALU
#define PMT_QUAD_COMBINE_ALU 1
#endif
__device__ __forceinline__ void permute_quad(uint64_t (&e)[4][3], const poseidonx::QuadTables& T, const poseidonx::QuadFrags& f,
                                             unsigned lane) {
  poseidonx::permute_quad<PMT_QUAD_SBOX_FMA_MASK, PMT_QUAD_PART_FMA_MASK, PMT_QUAD_COMBINE_ALU != 0>(e, T, f, lane);
}


template <class Layout>
__global__ void __launch_bounds__(BLOCK, PMT_MINB) k_level(Layout lay, int l, size_t k0, size_t count) {
#if PMT_QUAD
  __shared__ poseidonx::QuadTables T;
  poseidonx::quad_stage_tables(T);
  const unsigned lane = threadIdx.x & 31, q = lane >> 2, j = lane & 3;
  poseidonx::QuadFrags f;
  poseidonx::quad_load_frags(f, q, j);
#if PMT_QUAD_FRAGS_SMEM
  poseidonx::quad_publish_frags(T, f, lane);
#endif
  for (size_t base = (size_t)blockIdx.x * BLOCK; base < count; base += (size_t)gridDim.x * BLOCK) {
    const size_t wbase = base + (threadIdx.x & ~31u);
    if (wbase >= count) continue;                       // warp-uniform
    uint64_t e[4][3];
#pragma unroll
    for (int mb = 0; mb < 4; mb++) {
      size_t i = wbase + 8 * mb + q;
      if (i >= count) i = count - 1;                    // ragged tail: recompute the last node, store nothing
      const uint64_t *a, *b;
      lay.children(l, k0 + i, a, b);
      e[mb][0] = a[j]; e[mb][1] = b[j]; e[mb][2] = 0;
    }
    permute_quad(e, T, f, lane);
#pragma unroll
    for (int mb = 0; mb < 4; mb++) {
      const size_t i = wbase + 8 * mb + q;
      if (i < count) lay.at(l, k0 + i)[j] = gl::canonical(e[mb][0]);
    }
  }
#else
  for (size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x; i < count; i += (size_t)gridDim.x * BLOCK) {
    const uint64_t *a, *b;
    lay.children(l, k0 + i, a, b);
    store_digest(lay.at(l, k0 + i), two_to_one_ab(load_digest(a), load_digest(b)));
  }
#endif
}
}  // namespace ab

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void k_fill(uint64_t* p, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint64_t z = (i + 1) * 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; z ^= z >> 31;
    p[i] = (i % 1000 == 7) ? ~0ull - (i & 3) : z;     // a few non-canonical inputs
  }
}
__global__ void k_sum(const uint64_t* p, size_t n, unsigned long long* out) {
  unsigned long long a = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    a += p[i] * (2 * i + 1);
  atomicAdd(out, a);
}

int main(int argc, char** argv) {
  const char* name = argc > 1 ? argv[1] : "variant";
  const int lg = argc > 2 ? atoi(argv[2]) : 23;          // leaves
  const int blocks_per_sm = argc > 3 ? atoi(argv[3]) : 0; // 0 = occupancy query
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const size_t n = (size_t)1 << lg;
  uint64_t *dig, *cap; unsigned long long* dsum;
  CK(cudaMalloc(&dig, (2 * n - 2) * 32)); CK(cudaMalloc(&cap, 32)); CK(cudaMalloc(&dsum, 8));
  pmt::Plonky2 lay{dig, cap, lg};
  k_fill<<<1024, 256>>>(dig, (2 * n - 2) * 4);
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ab::k_level<pmt::Plonky2>, pmt::BLOCK, 0));
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, ab::k_level<pmt::Plonky2>));
  const int bps = blocks_per_sm ? blocks_per_sm : occ;
  // blocks_per_sm < 0: one node per thread (what pmt_api.cu launches); otherwise a persistent grid of resident blocks
  const unsigned grid = blocks_per_sm < 0 ? (unsigned)((n / 2 + pmt::BLOCK - 1) / pmt::BLOCK) : (unsigned)(prop.multiProcessorCount * bps);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const size_t count = n / 2;
  float best = 1e30f;
  for (int rep = 0; rep < 6; rep++) {
    CK(cudaEventRecord(e0));
    ab::k_level<pmt::Plonky2><<<grid, pmt::BLOCK>>>(lay, 1, 0, count);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  // level 2 as well, so the checksum covers outputs that were inputs
  ab::k_level<pmt::Plonky2><<<grid, pmt::BLOCK>>>(lay, 2, 0, count / 2);
  CK(cudaMemset(dsum, 0, 8));
  k_sum<<<1024, 256>>>(dig, (2 * n - 2) * 4, dsum);
  unsigned long long h; CK(cudaMemcpy(&h, dsum, 8, cudaMemcpyDeviceToHost));
  printf("{\"ab\": \"%s\", \"regs\": %d, \"local_bytes\": %zu, \"occ_blocks\": %d, \"block\": %d, \"grid\": %u, \"nodes\": %zu, \"ms\": %.4f, \"Gperm_per_s\": %.4f, \"checksum\": \"%016llx\"}\n",
         name, fa.numRegs, (size_t)fa.localSizeBytes, occ, pmt::BLOCK, grid, count, best, count / best * 1e-6, h);
  return 0;
}
