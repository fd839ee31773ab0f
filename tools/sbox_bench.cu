// sbox_bench.cu -- tuning harness (not part of the product library): throughput of the Goldilocks S-box x^7 (4 modular
// multiplications) and of the fp64 -> u64 recombination alone, at several occupancies, to find what the integer part of
// the permutation costs on B200 when nothing else competes.  Prints one JSON object per line.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "experimental/poseidon_variants.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// MODE 0: pow7_mix<MASK> on LANES independent lanes   1: combine_magic_alu   2: combine_magic_fma
// 3: mul only (glx::mul<ALU = !(MASK & 1)>)
template <int MODE, int MASK, int LANES>
__global__ void __launch_bounds__(128) k_sbox(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, int iters) {
  uint64_t s[LANES];
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int i = 0; i < LANES; i++) s[i] = in[(t * LANES + i) & 4095];
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < LANES; i++) {
      if (MODE == 0) s[i] = poseidonx::pow7_mix<MASK>(s[i]);
      if (MODE == 1) s[i] = poseidonx::combine_magic_alu(__longlong_as_double((long long)(0x4330000000000000ull | (s[i] >> 14))),
                                                         __longlong_as_double((long long)(0x4330000000000000ull | (s[(i + 1) % LANES] >> 13))));
      if (MODE == 2) s[i] = poseidonx::combine_magic_fma(__longlong_as_double((long long)(0x4330000000000000ull | (s[i] >> 14))),
                                                         __longlong_as_double((long long)(0x4330000000000000ull | (s[(i + 1) % LANES] >> 13))));
      if (MODE == 3) s[i] = (MASK & 1) ? glx::mul<false>(s[i], s[(i + 1) % LANES]) : glx::mul<true>(s[i], s[(i + 1) % LANES]);
    }
  }
  uint64_t r = 0;
#pragma unroll
  for (int i = 0; i < LANES; i++) r ^= s[i];
  out[t] = r;
}

template <int MODE, int MASK, int LANES>
static void run(const char* name, int sms, int warps_per_sm, double ops_per_lane_iter) {
  const int threads = 128, blocks = sms * warps_per_sm / 4, iters = 256;
  uint64_t *din, *dout; CK(cudaMalloc(&din, 4096 * 8)); CK(cudaMalloc(&dout, (size_t)threads * blocks * 8));
  CK(cudaMemset(din, 0x5a, 4096 * 8));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_sbox<MODE, MASK, LANES><<<blocks, threads>>>(din, dout, 4);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(e0));
    k_sbox<MODE, MASK, LANES><<<blocks, threads>>>(din, dout, iters);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k_sbox<MODE, MASK, LANES>));
  const double ops_per_warp = (double)iters * LANES * ops_per_lane_iter;          // per warp
  const double cyc = best * 1e-3 * 1.965e9 / (ops_per_warp * warps_per_sm / 4);   // cycles per op per SMSP
  printf("{\"bench\": \"sbox\", \"name\": \"%s\", \"lanes\": %d, \"warps_per_sm\": %d, \"regs\": %d, \"ms\": %.4f, \"cycles_per_op_per_smsp\": %.2f}\n",
         name, LANES, warps_per_sm, fa.numRegs, best, cyc);
  CK(cudaFree(din)); CK(cudaFree(dout));
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  for (int w : {4, 8, 12, 16, 32}) {
    run<0, 0, 12>("pow7 alu-reduce (per mul)", sms, w, 4);
    run<0, 15, 12>("pow7 fma-reduce (per mul)", sms, w, 4);
    run<0, 5, 12>("pow7 mask5 (per mul)", sms, w, 4);
  }
  for (int w : {12, 32}) {
    run<0, 0, 3>("pow7 alu-reduce 3 lanes (per mul)", sms, w, 4);
    run<3, 0, 12>("mul alu-reduce", sms, w, 1);
    run<3, 1, 12>("mul fma-reduce", sms, w, 1);
    run<1, 0, 12>("combine_magic_alu", sms, w, 1);
    run<2, 0, 12>("combine_magic_fma", sms, w, 1);
  }
  return 0;
}
