// dmma_bench.cu -- tuning harness (not part of the product library): throughput of the fp64 tensor-core MMA (DMMA) on
// B200 alone and mixed with ALU / IMAD work, to decide whether the Poseidon MDS layer (a 12 x 12 matrix times a batch of
// states) should move from DFMA to DMMA.  Prints one JSON object per line.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o dmma_bench tools/dmma_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&d)[4], double a0, double a1, double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a0), "d"(a1), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], double b0, double b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b0), "d"(b1));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// MODE 0: m8n8k4   1: m16n8k4   2: m16n8k8   3: m16n8k16
// MIX: per MMA, this many extra instructions of kind KIND (0 none, 1 LOP3, 2 IMAD.WIDE zero-addend, 3 DFMA, 4 IMAD lo)
template <int MODE, int KIND, int MIX>
__global__ void __launch_bounds__(256) k_dmma(double* out, double a_in, uint32_t u, int iters) {
  constexpr int ILP = 4;
  double acc[ILP][4];
  double a[8], b[4];
  uint32_t x[8]; uint64_t w[8]; double f[8];
#pragma unroll
  for (int j = 0; j < 8; j++) { a[j] = a_in + j + threadIdx.x; x[j] = threadIdx.x * 7 + j; w[j] = j; f[j] = 1.0 + j; }
#pragma unroll
  for (int j = 0; j < 4; j++) b[j] = a_in * 0.5 + j;
#pragma unroll
  for (int j = 0; j < ILP; j++) { acc[j][0] = j; acc[j][1] = 1; acc[j][2] = 2; acc[j][3] = 3; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int rep = 0; rep < 4; rep++) {
#pragma unroll
      for (int j = 0; j < ILP; j++) {
        if (MODE == 0) dmma884(acc[j][0], acc[j][1], a[j], b[0]);
        if (MODE == 1) dmma1684(acc[j], a[j], a[j + 1], b[0]);
        if (MODE == 2) { double aa[4] = {a[0], a[1], a[2], a[3]}; dmma1688(acc[j], aa, b[0], b[1]); }
        if (MODE == 3) dmma16816(acc[j], a, b);
#pragma unroll
        for (int m = 0; m < MIX; m++) {
          const int t = (j * MIX + m) % 8;
          if (KIND == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[t]) : "r"(x[(t + 1) % 8]), "r"(u));
          if (KIND == 2) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[t]) : "r"(x[t]), "r"((uint32_t)w[(t + 1) % 8]));
          if (KIND == 3) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f[t]) : "d"(a_in), "d"(f[(t + 1) % 8]));
          if (KIND == 4) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[t]) : "r"(u), "r"(x[(t + 1) % 8]));
        }
      }
    }
  }
  double r = 0;
#pragma unroll
  for (int j = 0; j < ILP; j++) r += acc[j][0] + acc[j][1] + acc[j][2] + acc[j][3];
#pragma unroll
  for (int j = 0; j < 8; j++) r += (double)x[j] + (double)w[j] + f[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE, int KIND, int MIX>
static void run(const char* name, int sms, int warps_per_sm) {
  const int threads = 256, blocks = sms * warps_per_sm / 8, iters = 2048;
  double* d; CK(cudaMalloc(&d, sizeof(double) * threads * blocks));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_dmma<MODE, KIND, MIX><<<blocks, threads>>>(d, 1.25, 77u, 8);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(e0));
    k_dmma<MODE, KIND, MIX><<<blocks, threads>>>(d, 1.25, 77u, iters);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  const double macs_per_mma = MODE == 0 ? 256 : MODE == 1 ? 512 : MODE == 2 ? 1024 : 2048;
  const double mmas = (double)blocks * (threads / 32) * iters * 16;
  const double cyc_per_mma_smsp = best * 1e-3 * 1.965e9 / ((double)iters * 16 * warps_per_sm / 4);
  printf("{\"bench\": \"dmma\", \"name\": \"%s\", \"warps_per_sm\": %d, \"ms\": %.4f, \"TMAC_per_s\": %.2f, \"cycles_per_mma_per_smsp\": %.2f, "
         "\"extra_instr_per_mma\": %d}\n", name, warps_per_sm, best, mmas * macs_per_mma / best * 1e-9, cyc_per_mma_smsp, MIX);
  CK(cudaFree(d));
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  for (int w : {8, 16, 32}) {
    run<0, 0, 0>("m8n8k4", sms, w);
    run<1, 0, 0>("m16n8k4", sms, w);
    run<2, 0, 0>("m16n8k8", sms, w);
    run<3, 0, 0>("m16n8k16", sms, w);
  }
  run<0, 1, 4>("m8n8k4 + 4 lop3", sms, 16);
  run<0, 1, 8>("m8n8k4 + 8 lop3", sms, 16);
  run<0, 1, 16>("m8n8k4 + 16 lop3", sms, 16);
  run<0, 2, 4>("m8n8k4 + 4 imad.wide", sms, 16);
  run<0, 2, 8>("m8n8k4 + 8 imad.wide", sms, 16);
  run<0, 3, 4>("m8n8k4 + 4 dfma", sms, 16);
  run<0, 3, 8>("m8n8k4 + 8 dfma", sms, 16);
  run<0, 4, 8>("m8n8k4 + 8 imad.lo", sms, 16);
  run<1, 1, 8>("m16n8k4 + 8 lop3", sms, 16);
  run<1, 1, 16>("m16n8k4 + 16 lop3", sms, 16);
  run<1, 2, 8>("m16n8k4 + 8 imad.wide", sms, 16);
  run<3, 1, 16>("m16n8k16 + 16 lop3", sms, 16);
  run<3, 1, 32>("m16n8k16 + 32 lop3", sms, 16);
  return 0;
}
