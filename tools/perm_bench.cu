// perm_bench.cu -- tuning harness (not part of the product library): integer-pipe peak microbenchmarks and
// raw permutation throughput on one GPU.  Build: see tools/build_perm_bench.sh.  Prints one JSON object per line.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#include "experimental/poseidon_variants.cuh"
#include "../plonky2_merkle_trees_b200/csrc/poseidon_coop.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// ---------------------------------------------------------------------------------------------------------------
// pipe peaks: ILP independent chains per thread, ITER iterations, counted ops = threads * ILP * ITER
// ---------------------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) k_pipe(uint64_t* out, uint32_t a, uint32_t b, int iters) {
  constexpr int ILP = 8;
  uint64_t acc[ILP];
  uint32_t x[ILP], y[ILP], z[ILP];
  double dx[ILP];
  const double dk = (double)a * 1e-9;
#pragma unroll
  for (int j = 0; j < ILP; j++) { acc[j] = threadIdx.x + j; x[j] = threadIdx.x * 7 + j; y[j] = x[j] + 1; z[j] = j; dx[j] = 1.0 + j + threadIdx.x; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int j = 0; j < ILP; j++) {
        if (MODE == 0) {  // IMAD.WIDE.U32 acc += a * b'
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(a), "r"(x[j]));
        } else if (MODE == 1) {  // IMAD (32-bit lo)
          asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[j]) : "r"(a), "r"(b));
        } else if (MODE == 2) {  // IADD3
          asm volatile("add.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(a));
        } else if (MODE == 3) {  // LOP3
          asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[j]) : "r"(a));
        } else if (MODE == 4) {  // IMAD.WIDE + IADD3 interleaved (both counted)
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(a), "r"(b));
          asm volatile("add.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(a));
        } else if (MODE == 5) {  // carry chain: add.cc / addc (IADD3 + IADD3.X), counted as 2
          uint32_t lo = (uint32_t)acc[j], hi = (uint32_t)(acc[j] >> 32);
          asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
          acc[j] = ((uint64_t)hi << 32) | lo;
        } else if (MODE == 6) {  // IMAD.WIDE + 2x IADD3 (1:2 mix)
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(a), "r"(b));
          asm volatile("add.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(a));
          asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[j]) : "r"(b));
        } else if (MODE == 7) {  // SHF (funnel shift)
          asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(x[j]) : "r"(a));
        } else if (MODE == 8) {  // IADD3, operands that cannot be folded
          asm volatile("add.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(x[(j + 1) % ILP]));
        } else if (MODE == 9) {  // IMAD.HI.U32 with accumulate
          asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(x[j]) : "r"(a), "r"(x[(j + 1) % ILP]));
        } else if (MODE == 10) {  // IMAD.WIDE.U32 with carry-out + carry accumulate (counted as 2)
          uint32_t lo = (uint32_t)acc[j], hi = (uint32_t)(acc[j] >> 32);
          asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                       : "+r"(lo), "+r"(hi), "+r"(x[j]) : "r"(a), "r"(x[(j + 1) % ILP]));
          acc[j] = ((uint64_t)hi << 32) | lo;
        } else if (MODE == 11) {  // chained IMAD.WIDE.U32 with immediate coefficient (the MDS form)
          uint32_t lo = (uint32_t)acc[j], hi = (uint32_t)(acc[j] >> 32);
          asm volatile("mad.lo.cc.u32 %0, %2, 41, %0;\n\tmadc.hi.u32 %1, %2, 41, %1;" : "+r"(lo), "+r"(hi) : "r"(x[j]));
          acc[j] = ((uint64_t)hi << 32) | lo;
        } else if (MODE == 12) {  // LOP3 unfoldable
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(x[(j + 1) % ILP]), "r"(a));
        } else if (MODE == 14) {  // DFMA (fp64 pipe), dependent accumulate
          double d = __longlong_as_double((long long)acc[j]);
          asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d) : "d"(dk), "d"(dx[j]));
          acc[j] = (uint64_t)__double_as_longlong(d);
        } else if (MODE == 15) {  // DFMA + chained IMAD.WIDE (accumulate form) in parallel (counted as 2)
          double d = __longlong_as_double((long long)acc[j]);
          asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d) : "d"(dk), "d"(dx[j]));
          acc[j] = (uint64_t)__double_as_longlong(d);
          asm volatile("mad.lo.cc.u32 %0, %2, 41, %0;\n\tmadc.hi.u32 %1, %2, 41, %1;" : "+r"(y[j]), "+r"(z[j]) : "r"(x[j]));
        } else if (MODE == 16) {  // DFMA + IMAD.WIDE acc + IADD3 (counted as 3)
          double d = __longlong_as_double((long long)acc[j]);
          asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d) : "d"(dk), "d"(dx[j]));
          acc[j] = (uint64_t)__double_as_longlong(d);
          asm volatile("mad.lo.cc.u32 %0, %2, 41, %0;\n\tmadc.hi.u32 %1, %2, 41, %1;" : "+r"(y[j]), "+r"(z[j]) : "r"(x[j]));
          asm volatile("add.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(x[(j + 1) % ILP]));
        } else if (MODE == 17) {  // DADD
          double d = __longlong_as_double((long long)acc[j]);
          asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d) : "d"(dk));
          acc[j] = (uint64_t)__double_as_longlong(d);
        } else if (MODE == 18) {  // DFMA + IADD3 (1:1): does an fp64 instruction hold the issue port for 2 cycles?
          double d = __longlong_as_double((long long)acc[j]);
          asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d) : "d"(dk), "d"(dx[j]));
          acc[j] = (uint64_t)__double_as_longlong(d);
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(x[(j + 1) % ILP]), "r"(a));
        } else if (MODE == 19) {  // DFMA + 2 ALU (1:2)
          double d = __longlong_as_double((long long)acc[j]);
          asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d) : "d"(dk), "d"(dx[j]));
          acc[j] = (uint64_t)__double_as_longlong(d);
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(x[(j + 1) % ILP]), "r"(a));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[j]) : "r"(y[(j + 1) % ILP]), "r"(a));
        } else if (MODE == 20) {  // I2F.F64.U32
          asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(dx[j]) : "r"(x[j]));
          x[j] += (uint32_t)__double2hiint(dx[j]);
        } else if (MODE == 21) {  // DFMA + IMAD lo (1:1): same pipe or not
          double d = __longlong_as_double((long long)acc[j]);
          asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d) : "d"(dk), "d"(dx[j]));
          acc[j] = (uint64_t)__double_as_longlong(d);
          asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[j]) : "r"(a), "r"(b));
        } else if (MODE == 22) {  // DFMA + IMAD lo + LOP3 (1:1:1)
          double d = __longlong_as_double((long long)acc[j]);
          asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d) : "d"(dk), "d"(dx[j]));
          acc[j] = (uint64_t)__double_as_longlong(d);
          asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[j]) : "r"(a), "r"(b));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[j]) : "r"(y[(j + 1) % ILP]), "r"(a));
        } else if (MODE == 23) {  // pure-ALU carry chain with a carry-in AND carry-out in the middle (sub.cc / subc.cc / subc)
          asm volatile("sub.cc.u32 %0, %0, %3;\n\tsubc.cc.u32 %1, %1, %4;\n\tsubc.u32 %2, %2, 0;"
                       : "+r"(x[j]), "+r"(y[j]), "+r"(z[j]) : "r"(x[(j + 1) % ILP]), "r"(y[(j + 1) % ILP]));
        } else if (MODE == 24) {  // 2 DFMA + 1 LOP3
          double d = __longlong_as_double((long long)acc[j]);
          asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d) : "d"(dk), "d"(dx[j]));
          asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(dx[j]) : "d"(dk), "d"(dk));
          acc[j] = (uint64_t)__double_as_longlong(d);
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(x[(j + 1) % ILP]), "r"(a));
        } else if (MODE == 30) {  // IMAD.WIDE.U32 zero addend, operands depend on the previous result (nothing to hoist)
          asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(acc[j]) : "r"((uint32_t)acc[j]), "r"(y[j]));
        } else if (MODE == 31) {  // IMAD.WIDE.U32 accumulate form, multiplicand depends on the accumulator
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"((uint32_t)acc[j]), "r"(y[j]));
        } else if (MODE == 32) {  // zero-addend IMAD.WIDE + LOP3 (1:1)
          asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(acc[j]) : "r"((uint32_t)acc[j]), "r"(y[j]));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(x[(j + 1) % ILP]), "r"(a));
        } else if (MODE == 33) {  // accumulate-form IMAD.WIDE + LOP3 (1:1)
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"((uint32_t)acc[j]), "r"(y[j]));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(x[(j + 1) % ILP]), "r"(a));
        } else if (MODE == 34) {  // accumulate-form IMAD.WIDE + 2 LOP3 (1:2)
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"((uint32_t)acc[j]), "r"(y[j]));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(x[(j + 1) % ILP]), "r"(a));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[j]) : "r"(z[(j + 1) % ILP]), "r"(a));
        } else if (MODE == 35) {  // zero-addend IMAD.WIDE + carry chain add.cc/addc (1:2)
          asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(acc[j]) : "r"((uint32_t)acc[j]), "r"(y[j]));
          asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(x[j]), "+r"(z[j]) : "r"(x[(j + 1) % ILP]), "r"(z[(j + 1) % ILP]));
        } else if (MODE == 36) {  // 32-bit IMAD with data-dependent operands + LOP3 (1:1)
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[j]) : "r"(y[(j + 1) % ILP]), "r"(a));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(x[(j + 1) % ILP]), "r"(a));
        } else if (MODE == 37) {  // 32-bit IMAD alone, data-dependent
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[j]) : "r"(y[(j + 1) % ILP]), "r"(a));
        } else if (MODE == 38) {  // IMAD.WIDE.U32 zero addend with BOTH result words consumed (so it stays an IMAD.WIDE) + 1 LOP3
          uint64_t w;
          asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(x[j]), "r"(y[j]));
          asm volatile("lop3.b32 %0, %1, %2, %0, 0x96;" : "+r"(x[j]) : "r"((uint32_t)w), "r"((uint32_t)(w >> 32)));
        } else if (MODE == 40) {  // FFMA, three register operands
          float f = __uint_as_float(x[j]);
          asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(__uint_as_float(y[(j + 1) % ILP])), "f"(__uint_as_float(z[j])));
          x[j] = __float_as_uint(f);
        } else if (MODE == 41) {  // FFMA2 (fma.rn.f32x2, Blackwell packed fp32): counted as ONE instruction = 2 fp32 FMAs
          asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[j]) : "l"(acc[(j + 1) % ILP]), "l"((uint64_t)a << 32 | b));
        } else if (MODE == 42) {  // FFMA2 + LOP3 (1:1): does the packed form leave the alu pipe free?
          asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[j]) : "l"(acc[(j + 1) % ILP]), "l"((uint64_t)a << 32 | b));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(x[(j + 1) % ILP]), "r"(a));
        } else if (MODE == 43) {  // FFMA2 + IMAD.WIDE accumulate (1:1): shared pipe or not
          asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[j]) : "l"(acc[(j + 1) % ILP]), "l"((uint64_t)a << 32 | b));
          uint64_t w = (uint64_t)z[j] << 32 | y[j];
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w) : "r"(x[j]), "r"(a));
          y[j] = (uint32_t)w; z[j] = (uint32_t)(w >> 32);
        } else if (MODE == 44) {  // FADD2 (add.rn.f32x2)
          asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(acc[j]) : "l"(acc[(j + 1) % ILP]));
        } else if (MODE == 13) {  // SEL
          asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %1, %2, p; }" : "+r"(x[j]) : "r"(x[(j + 1) % ILP]), "r"(a));
        }
      }
    }
  }
  uint64_t r = 0;
#pragma unroll
  for (int j = 0; j < ILP; j++) r += acc[j] + x[j] + y[j] + z[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// ---------------------------------------------------------------------------------------------------------------
// permutation throughput: each thread runs `chain` dependent permutations on its own state
// ---------------------------------------------------------------------------------------------------------------
#ifndef MINB
#define MINB 1
#endif
template <int VARIANT>
__global__ void __launch_bounds__(128, MINB) k_perm(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, int chain) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t s[12];
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = in[t * 12 + i];
  for (int c = 0; c < chain; c++) {
    if (VARIANT == 0) poseidonx::permute_naive<false>(s);
    if (VARIANT == 1) poseidonx::permute_fast<false, false>(s);
    if (VARIANT == 2) poseidonx::permute_fast<true, false>(s);
    if (VARIANT == 3) poseidonx::permute_fast<true, true>(s);
    if (VARIANT == 4) poseidonx::permute_naive<true>(s);
    if (VARIANT == 5) poseidonx::permute_fast<true, true, 1>(s);
    if (VARIANT == 6) poseidonx::permute_fast<false, false, 1>(s);
    if (VARIANT == 7) poseidonx::permute_fast<false, false, 2>(s);
    if (VARIANT == 8) poseidonx::permute_fast<true, true, 2>(s);
    if (VARIANT == 10) poseidonx::permute_fast<true, true, 2, false, false, true>(s);
    if (VARIANT == 11) poseidonx::permute_fast<true, true, 3>(s);
    if (VARIANT == 12) poseidonx::permute_fast<false, false, 3>(s);
    if (VARIANT == 9) { s[8] = s[9] = s[10] = s[11] = 0; poseidonx::permute_fast<true, true, 2, true, true>(s); }
  }
#pragma unroll
  for (int i = 0; i < 12; i++) out[t * 12 + i] = glx::canonical(s[i]);
}

// single-warp latency of the cooperative (16 lanes per state) permutation
__global__ void __launch_bounds__(256) k_perm_coop(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, int chain) {
  __shared__ uint64_t rc_smem[12 * 31];
  for (int i = threadIdx.x; i < 12 * 31; i += blockDim.x) rc_smem[i] = PMT_RC[i];
  __syncthreads();
  const unsigned g = threadIdx.x & 15, base_lane = threadIdx.x & 16;
  const size_t grp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  uint64_t v = g < 12 ? in[grp * 12 + g] : 0;
  for (int c = 0; c < chain; c++) v = poseidonx::permute_coop(v, rc_smem, g, base_lane);
  if (g < 12) out[grp * 12 + g] = glx::canonical(v);
}

// the product's cooperative forms (csrc/poseidon_coop.cuh): Quad = four threads per state, eight states per warp;
// Wide = 16 lanes per state, two states per warp
template <class Form>
__global__ void __launch_bounds__(256) k_perm_form(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, int chain) {
  __shared__ poseidon::coop::Shared<8> sh;
  const Form t = Form::make(sh);
  poseidon::coop::stage(sh);
  const size_t st = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / Form::LANES;
  uint64_t e[Form::ELEMS];
  for (int a = 0; a < Form::ELEMS; a++) e[a] = t.elem(a) < 12 ? in[st * 12 + t.elem(a)] : 0;
  for (int c = 0; c < chain; c++) t.permute(e, sh);
  for (int a = 0; a < Form::ELEMS; a++) if (t.elem(a) < 12) out[st * 12 + t.elem(a)] = gl::canonical(e[a]);
}
// the product's thread-per-state form
__global__ void __launch_bounds__(128) k_perm_prod(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, int chain) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t s[12];
  for (int i = 0; i < 12; i++) s[i] = in[t * 12 + i];
  for (int c = 0; c < chain; c++) poseidon::permute(s);
  for (int i = 0; i < 12; i++) out[t * 12 + i] = gl::canonical(s[i]);
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

template <int MODE>
static void run_pipe(const char* name, double ops_per_inner, int sms) {
  const int threads = 256, blocks = sms * 8, iters = 4096;
  uint64_t* d; CK(cudaMalloc(&d, sizeof(uint64_t) * threads * blocks));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_pipe<MODE><<<blocks, threads>>>(d, 12345u, 6789u, 16);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    CK(cudaEventRecord(e0));
    k_pipe<MODE><<<blocks, threads>>>(d, 12345u, 6789u, iters);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms = time_ms(e0, e1); if (ms < best) best = ms;
  }
  double ops = (double)threads * blocks * 8 * 4 * iters * ops_per_inner;
  printf("{\"bench\": \"pipe\", \"name\": \"%s\", \"ms\": %.4f, \"Gops_per_s\": %.1f, \"ops_per_clk_per_sm_at_1965MHz\": %.2f}\n",
         name, best, ops / best * 1e-6, ops / (best * 1e-3) / 1.965e9 / sms);
  CK(cudaFree(d));
}

template <int VARIANT>
static void run_perm(const char* name, int sms, int blocks_per_sm, int chain, bool check) {
  const int threads = 128;
  const size_t n = (size_t)threads * sms * blocks_per_sm;
  std::vector<uint64_t> h(n * 12);
  for (size_t i = 0; i < n * 12; i++) h[i] = i < 12 ? i : (i * 0x9E3779B97F4A7C15ull) ^ (i >> 7);
  uint64_t *din, *dout; CK(cudaMalloc(&din, n * 96)); CK(cudaMalloc(&dout, n * 96));
  CK(cudaMemcpy(din, h.data(), n * 96, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_perm<VARIANT><<<(unsigned)(n / threads), threads>>>(din, dout, 1);
  CK(cudaDeviceSynchronize());
  if (check) {
    uint64_t r[4]; CK(cudaMemcpy(r, dout, 32, cudaMemcpyDeviceToHost));
    const uint64_t want[4] = {0xd64e1e3efc5b8e9eull, 0x53666633020aaa47ull, 0xd40285597c6a8825ull, 0x613a4f81e81231d2ull};
    printf("{\"bench\": \"kat\", \"name\": \"%s\", \"ok\": %s, \"got0\": \"%016llx\"}\n", name,
           memcmp(r, want, 32) == 0 ? "true" : "false", (unsigned long long)r[0]);
  }
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    CK(cudaEventRecord(e0));
    k_perm<VARIANT><<<(unsigned)(n / threads), threads>>>(din, dout, chain);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms = time_ms(e0, e1); if (ms < best) best = ms;
  }
  printf("{\"bench\": \"perm\", \"name\": \"%s\", \"blocks_per_sm\": %d, \"threads\": %d, \"chain\": %d, \"ms\": %.4f, \"Gperm_per_s\": %.4f}\n",
         name, blocks_per_sm, threads, chain, best, (double)n * chain / best * 1e-6);
  CK(cudaFree(din)); CK(cudaFree(dout));
}

static void run_latency(int sms) {
  uint64_t *din, *dout; CK(cudaMalloc(&din, 1 << 23)); CK(cudaMalloc(&dout, 1 << 23));
  std::vector<uint64_t> h(1 << 20);
  for (size_t i = 0; i < h.size(); i++) h[i] = i;
  CK(cudaMemcpy(din, h.data(), 1 << 23, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int chain : {1, 17}) {
    for (int rep = 0; rep < 3; rep++) {
      CK(cudaEventRecord(e0)); k_perm_coop<<<1, 32>>>(din, dout, chain); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float a = time_ms(e0, e1);
      CK(cudaEventRecord(e0)); k_perm<8><<<1, 32>>>(din, dout, chain); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float b = time_ms(e0, e1);
      CK(cudaEventRecord(e0)); k_perm_coop<<<sms, 256>>>(din, dout, chain); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float c2 = time_ms(e0, e1);
      if (rep == 2) printf("{\"bench\": \"latency\", \"chain\": %d, \"coop_1warp_us\": %.2f, \"thread_1warp_us\": %.2f, \"coop_148x256_us\": %.2f}\n", chain, a * 1e3, b * 1e3, c2 * 1e3);
    }
  }
  uint64_t r[4]; CK(cudaMemcpy(r, dout, 32, cudaMemcpyDeviceToHost));
  poseidon::coop::k_coop_tables_init<<<1, 256>>>();      // the cooperative forms stage their constants from global memory
  CK(cudaDeviceSynchronize());
  // round 2: the Quad and Wide forms against round 1's 16-lane form and the product's thread-per-state form: chains of
  // dependent permutations on one warp / half a block / one block / one block per SM / three blocks per SM
  for (int chain : {1, 17}) {
    float q[5], w[5], tp = 0, l16 = 0;
    auto T = [&](auto launch) { CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); return time_ms(e0, e1); };
    for (int rep = 0; rep < 3; rep++) {
      q[0] = T([&] { k_perm_form<poseidon::coop::Quad><<<1, 32>>>(din, dout, chain); });
      q[1] = T([&] { k_perm_form<poseidon::coop::Quad><<<1, 128>>>(din, dout, chain); });
      q[2] = T([&] { k_perm_form<poseidon::coop::Quad><<<1, 256>>>(din, dout, chain); });
      q[3] = T([&] { k_perm_form<poseidon::coop::Quad><<<sms, 256>>>(din, dout, chain); });
      q[4] = T([&] { k_perm_form<poseidon::coop::Quad><<<3 * sms, 256>>>(din, dout, chain); });
      w[0] = T([&] { k_perm_form<poseidon::coop::Wide><<<1, 32>>>(din, dout, chain); });
      w[1] = T([&] { k_perm_form<poseidon::coop::Wide><<<1, 128>>>(din, dout, chain); });
      w[2] = T([&] { k_perm_form<poseidon::coop::Wide><<<1, 256>>>(din, dout, chain); });
      w[3] = T([&] { k_perm_form<poseidon::coop::Wide><<<sms, 256>>>(din, dout, chain); });
      w[4] = T([&] { k_perm_form<poseidon::coop::Wide><<<4 * sms, 256>>>(din, dout, chain); });
      tp = T([&] { k_perm_prod<<<1, 32>>>(din, dout, chain); });
      l16 = T([&] { k_perm_coop<<<1, 32>>>(din, dout, chain); });
    }
    printf("{\"bench\": \"latency_r2\", \"chain\": %d, \"quad_us\": {\"1x32\": %.2f, \"1x128\": %.2f, \"1x256\": %.2f, \"148x256\": %.2f, \"444x256\": %.2f}, "
           "\"wide_us\": {\"1x32\": %.2f, \"1x128\": %.2f, \"1x256\": %.2f, \"148x256\": %.2f, \"592x256\": %.2f}, \"thread_prod_1warp_us\": %.2f, \"lane16_r1_1warp_us\": %.2f}\n",
           chain, q[0] * 1e3, q[1] * 1e3, q[2] * 1e3, q[3] * 1e3, q[4] * 1e3, w[0] * 1e3, w[1] * 1e3, w[2] * 1e3, w[3] * 1e3, w[4] * 1e3, tp * 1e3, l16 * 1e3);
  }
  // KAT through both forms: perm(0..11)
  {
    std::vector<uint64_t> z(12 * 64);
    for (size_t i = 0; i < z.size(); i++) z[i] = i % 12;
    CK(cudaMemcpy(din, z.data(), z.size() * 8, cudaMemcpyHostToDevice));
    const uint64_t want[4] = {0xd64e1e3efc5b8e9eull, 0x53666633020aaa47ull, 0xd40285597c6a8825ull, 0x613a4f81e81231d2ull};
    k_perm_form<poseidon::coop::Quad><<<1, 256>>>(din, dout, 1);
    CK(cudaMemcpy(r, dout + 12 * 37, 32, cudaMemcpyDeviceToHost));
    printf("{\"bench\": \"kat\", \"name\": \"quad\", \"ok\": %s, \"got0\": \"%016llx\"}\n", memcmp(r, want, 32) == 0 ? "true" : "false", (unsigned long long)r[0]);
    k_perm_form<poseidon::coop::Wide><<<1, 256>>>(din, dout, 1);
    CK(cudaMemcpy(r, dout + 12 * 11, 32, cudaMemcpyDeviceToHost));
    printf("{\"bench\": \"kat\", \"name\": \"wide\", \"ok\": %s, \"got0\": \"%016llx\"}\n", memcmp(r, want, 32) == 0 ? "true" : "false", (unsigned long long)r[0]);
  }
  CK(cudaFree(din)); CK(cudaFree(dout));
}

int main(int argc, char** argv) {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("{\"bench\": \"device\", \"name\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);
  bool pipes = argc < 2 || strstr(argv[1], "pipe");
  bool perms = argc < 2 || strstr(argv[1], "perm");
  if (pipes) {
    run_pipe<0>("imad_wide_u32", 1, sms);
    run_pipe<1>("imad_lo_u32", 1, sms);
    run_pipe<2>("iadd3", 1, sms);
    run_pipe<3>("lop3", 1, sms);
    run_pipe<7>("shf", 1, sms);
    run_pipe<4>("imad_wide+iadd3 (1:1)", 2, sms);
    run_pipe<6>("imad_wide+2alu (1:2)", 3, sms);
    run_pipe<5>("iadd3.cc+iadd3.x", 2, sms);
    run_pipe<8>("iadd3_unfoldable", 1, sms);
    run_pipe<12>("lop3_unfoldable", 1, sms);
    run_pipe<13>("setp+selp", 2, sms);
    run_pipe<9>("imad_hi_u32", 1, sms);
    run_pipe<10>("imad_wide.cc+addc", 2, sms);
    run_pipe<11>("imad_wide_imm_chain", 1, sms);
    run_pipe<14>("dfma", 1, sms);
    run_pipe<17>("dadd", 1, sms);
    run_pipe<15>("dfma+imad_wide_acc (1:1)", 2, sms);
    run_pipe<16>("dfma+imad_wide_acc+iadd3 (1:1:1)", 3, sms);
    run_pipe<38>("imad_wide zero-addend, both words used (+1 lop3, counted 1)", 1, sms);
    run_pipe<30>("imad_wide zero-addend (dependent operands)", 1, sms);
    run_pipe<31>("imad_wide accumulate (dependent operands)", 1, sms);
    run_pipe<32>("imad_wide zero + lop3 (1:1)", 2, sms);
    run_pipe<33>("imad_wide acc + lop3 (1:1)", 2, sms);
    run_pipe<34>("imad_wide acc + 2 lop3 (1:2)", 3, sms);
    run_pipe<35>("imad_wide zero + add.cc/addc (1:2)", 3, sms);
    run_pipe<37>("imad_lo (dependent operands)", 1, sms);
    run_pipe<36>("imad_lo + lop3 (1:1)", 2, sms);
    run_pipe<18>("dfma+lop3 (1:1)", 2, sms);
    run_pipe<19>("dfma+2lop3 (1:2)", 3, sms);
    run_pipe<20>("i2f.f64.u32(+iadd)", 1, sms);
    run_pipe<21>("dfma+imad_lo (1:1)", 2, sms);
    run_pipe<22>("dfma+imad_lo+lop3 (1:1:1)", 3, sms);
    run_pipe<23>("sub.cc+subc.cc+subc", 3, sms);
    run_pipe<24>("2dfma+lop3 (2:1)", 3, sms);
  }
  if (argc > 1 && strstr(argv[1], "f32")) {   // fp32 candidates for a limb-form MDS layer (not used by the product kernels)
    run_pipe<40>("ffma", 1, sms);
    run_pipe<41>("ffma2 (f32x2, per instruction)", 1, sms);
    run_pipe<44>("fadd2 (f32x2, per instruction)", 1, sms);
    run_pipe<42>("ffma2+lop3 (1:1)", 2, sms);
    run_pipe<43>("ffma2+imad_wide_acc (1:1)", 2, sms);
    run_pipe<14>("dfma", 1, sms);
  }
  if (argc > 1 && strstr(argv[1], "lat")) run_latency(sms);
  if (perms) {
    int only_bps = argc > 2 ? atoi(argv[2]) : 0;
    for (int bps : {1, 2, 4, 8}) if (!only_bps || bps == only_bps) run_perm<0>("naive", sms, bps, 16, true);
    for (int bps : {2, 4, 8}) if (!only_bps || bps == only_bps) run_perm<1>("fast", sms, bps, 16, true);
    for (int bps : {2, 4, 8}) if (!only_bps || bps == only_bps) run_perm<2>("fast_sboxalu", sms, bps, 16, true);
    for (int bps : {2, 4, 8}) if (!only_bps || bps == only_bps) run_perm<3>("fast_allalu", sms, bps, 16, true);
    for (int bps : {2, 4, 8}) if (!only_bps || bps == only_bps) run_perm<5>("fast_allalu_limbmds", sms, bps, 16, true);
    for (int bps : {2, 4, 8}) if (!only_bps || bps == only_bps) run_perm<6>("fast_limbmds", sms, bps, 16, true);
    for (int bps : {2, 4, 8}) if (!only_bps || bps == only_bps) run_perm<7>("fast_dfmamds", sms, bps, 16, true);
    for (int bps : {2, 4, 8}) if (!only_bps || bps == only_bps) run_perm<8>("fast_allalu_dfmamds", sms, bps, 16, true);
    for (int bps : {2, 4, 8}) if (!only_bps || bps == only_bps) run_perm<11>("fast_allalu_dfmamds_i2f", sms, bps, 16, true);
    for (int bps : {2, 4, 8}) if (!only_bps || bps == only_bps) run_perm<12>("fast_dfmamds_i2f", sms, bps, 16, true);
    for (int bps : {2, 4, 8}) if (!only_bps || bps == only_bps) run_perm<10>("fast_allalu_dfmamds_sboxcall", sms, bps, 16, true);
    for (int bps : {2, 4, 8}) if (!only_bps || bps == only_bps) run_perm<9>("compress_allalu_dfmamds", sms, bps, 16, false);
  }
  return 0;
}
