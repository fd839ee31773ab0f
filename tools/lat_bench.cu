// lat_bench.cu -- tuning aid (not product code): dependent-chain latencies on one warp of a B200 SM, in cycles (clock64),
// of the building blocks of the cooperative permutations: x^7, a Goldilocks multiplication, the shuffle MDS row, DFMA /
// IMAD.WIDE / IADD3 chains, I2F, a shared-memory round trip.  One JSON line.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o lat_bench tools/lat_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../plonky2_merkle_trees_b200/csrc/poseidon_coop.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)
constexpr int N = 256;

template <int MODE>
__global__ void k_lat(uint64_t* out, long long* cycles, uint64_t seed) {
  __shared__ poseidon::coop::Shared<8> sh;
  __shared__ double2 buf[64];
  poseidon::coop::stage(sh);
  const poseidon::coop::Wide w = poseidon::coop::Wide::make(sh);
  uint64_t v = seed + threadIdx.x * 0x9e3779b97f4a7c15ull;
  double d = (double)(uint32_t)v;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; i++) {
    if (MODE == 0) v = gl::pow7(v);
    if (MODE == 1) v = gl::mul(v, v ^ 0x1234567);
    if (MODE == 2) v = gl::sqr(v);
    if (MODE == 3) { uint64_t L, H; w.row<false>(v, sh.rc[12 + w.gg], L, H); v = gl::combine_sums(L, H); }
    if (MODE == 4) { d = fma(d, 1.0000001, 3.0); d = fma(d, 1.0000001, 3.0); d = fma(d, 1.0000001, 3.0); d = fma(d, 1.0000001, 3.0); }   // 4 dependent DFMA
    if (MODE == 5) { v = gl::mad_wide(gl::lo32(v), 17u, v); v = gl::mad_wide(gl::lo32(v), 17u, v); v = gl::mad_wide(gl::lo32(v), 17u, v); v = gl::mad_wide(gl::lo32(v), 17u, v); }
    if (MODE == 6) { uint32_t a = gl::lo32(v); a = a * 3 + 1; a = (a ^ 5) + 7; a = a * 3 + 1; a = (a ^ 5) + 7; v = a; }   // mixed dependent int ops
    if (MODE == 7) { d = (double)(uint32_t)__double2loint(d) + 1.0; }     // I2F + DADD round trip
    if (MODE == 8) { buf[threadIdx.x] = make_double2(d, d); __syncwarp(); d = buf[(threadIdx.x + 1) & 31].x + 1.0; __syncwarp(); }
    if (MODE == 9) { v = poseidon::combine_magic(__longlong_as_double(0x4330000000000000ll | (long long)(v & 0xfffffffffffffll)), d); }
    if (MODE == 10) { uint32_t a = __shfl_sync(0xffffffffu, gl::lo32(v), (threadIdx.x + 1) & 31); v = a + 1; }
    if (MODE == 11) { v = gl::reduce128_c(gl::lo32(v), gl::hi32(v), gl::lo32(v) ^ 77u, gl::hi32(v) ^ 3u); }
  }
  long long t1 = clock64();
  if (MODE == 4 || MODE == 7 || MODE == 8) v ^= (uint64_t)__double_as_longlong(d);
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

template <int MODE>
static double run(uint64_t* d_out, long long* d_cyc, double per_iter_ops) {
  k_lat<MODE><<<1, 32>>>(d_out, d_cyc, 12345);
  k_lat<MODE><<<1, 32>>>(d_out, d_cyc, 12345);
  long long c = 0;
  cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
  return (double)c / N / per_iter_ops;
}

int main() {
  uint64_t* d_out; long long* d_cyc;
  CK(cudaMalloc(&d_out, 4096)); CK(cudaMalloc(&d_cyc, 64));
  poseidon::coop::k_coop_tables_init<<<1, 256>>>();
  CK(cudaDeviceSynchronize());
  printf("{\"bench\": \"chain_latency_cycles\", \"pow7\": %.1f, \"mul\": %.1f, \"sqr\": %.1f, \"shuffle_row_plus_combine\": %.1f, \"dfma\": %.1f, "
         "\"imad_wide_acc\": %.1f, \"int_alu_imad_mix_per_op\": %.1f, \"i2f_plus_dadd\": %.1f, \"sts_syncwarp_lds_dadd\": %.1f, \"combine_magic\": %.1f, "
         "\"shfl_plus_add\": %.1f, \"reduce128_c\": %.1f}\n",
         run<0>(d_out, d_cyc, 1), run<1>(d_out, d_cyc, 1), run<2>(d_out, d_cyc, 1), run<3>(d_out, d_cyc, 1), run<4>(d_out, d_cyc, 4), run<5>(d_out, d_cyc, 4),
         run<6>(d_out, d_cyc, 4), run<7>(d_out, d_cyc, 1), run<8>(d_out, d_cyc, 1), run<9>(d_out, d_cyc, 1), run<10>(d_out, d_cyc, 1), run<11>(d_out, d_cyc, 1));
  CK(cudaDeviceSynchronize());
  return 0;
}
