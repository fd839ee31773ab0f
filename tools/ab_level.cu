// ab_level.cu -- tuning harness (not part of the product library): times k_level<Plonky2> on one big level
// (2^22 two_to_one, inputs larger than L2) for whatever variant the -D flags select, and prints a checksum of the
// produced digests so that variants can be compared bit for bit.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -std=c++17 [-DPMT_...] -o ab_level tools/ab_level.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../plonky2_merkle_trees_b200/csrc/merkle_kernels.cuh"

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
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pmt::k_level<pmt::Plonky2>, pmt::BLOCK, 0));
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, pmt::k_level<pmt::Plonky2>));
  const int bps = blocks_per_sm ? blocks_per_sm : occ;
  // blocks_per_sm < 0: one node per thread (what pmt_api.cu launches); otherwise a persistent grid of resident blocks
  const unsigned grid = blocks_per_sm < 0 ? (unsigned)((n / 2 + pmt::BLOCK - 1) / pmt::BLOCK) : (unsigned)(prop.multiProcessorCount * bps);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const size_t count = n / 2;
  float best = 1e30f;
  for (int rep = 0; rep < 6; rep++) {
    CK(cudaEventRecord(e0));
    pmt::k_level<pmt::Plonky2><<<grid, pmt::BLOCK>>>(lay, 1, 0, count);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  // level 2 as well, so the checksum covers outputs that were inputs
  pmt::k_level<pmt::Plonky2><<<grid, pmt::BLOCK>>>(lay, 2, 0, count / 2);
  CK(cudaMemset(dsum, 0, 8));
  k_sum<<<1024, 256>>>(dig, (2 * n - 2) * 4, dsum);
  unsigned long long h; CK(cudaMemcpy(&h, dsum, 8, cudaMemcpyDeviceToHost));
  printf("{\"ab\": \"%s\", \"regs\": %d, \"local_bytes\": %zu, \"occ_blocks\": %d, \"block\": %d, \"grid\": %u, \"nodes\": %zu, \"ms\": %.4f, \"Gperm_per_s\": %.4f, \"checksum\": \"%016llx\"}\n",
         name, fa.numRegs, (size_t)fa.localSizeBytes, occ, pmt::BLOCK, grid, count, best, count / best * 1e-6, h);
  return 0;
}
