// Throughput of the SFU (MUFU) flavours the softmax / SiLU epilogues could use, ops per clock per SM (148 CTAs).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/mufu_peak2 tools/mufu_peak2.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
template <int MODE>
__global__ void k(float* out, int iters, long long* cycles) {
  float a[8];
  uint32_t u[8];
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3f + i; u[i] = 0x3c003c00u + threadIdx.x + i; }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
      if (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
      if (MODE == 3) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 4) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(u[i]));
      if (MODE == 5) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 6) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));        // F2FP
      if (MODE == 7) {                                   // ex2 and F2FP interleaved: do they share a pipe?
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
      }
      if (MODE == 8) {                                   // integer round-to-nearest pack: 2 x IADD + PRMT
        uint32_t lo = __float_as_uint(a[i]) + 0x8000u, hi = __float_as_uint(a[(i + 1) & 7]) + 0x8000u;
        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(u[i]) : "r"(lo), "r"(hi));
      }
      if (MODE == 9) {                                   // ex2 + integer pack
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        uint32_t lo = __float_as_uint(a[i]) + 0x8000u, hi = __float_as_uint(a[(i + 1) & 7]) + 0x8000u;
        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(u[i]) : "r"(lo), "r"(hi));
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = a[i] * 0.5f - 1.0f; u[i] ^= 0x00010001u; }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int MODE>
void run(const char* name, float* out, long long* cyc) {
  for (int threads : {256, 512, 1024}) {
    const int iters = 4096;
    k<MODE><<<148, threads>>>(out, iters, cyc); cudaDeviceSynchronize();
    k<MODE><<<148, threads>>>(out, iters, cyc); cudaDeviceSynchronize();
    printf("%-22s threads/SM %4d: %6.2f instr-lanes/clk/SM (x2 results for the packed forms)\n", name, threads,
           (double)threads * iters * 8 / (double)*cyc);
  }
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMallocManaged(&cyc, 8);
  run<0>("ex2.approx.ftz.f32", out, cyc);
  run<1>("ex2.approx.ftz.bf16x2", out, cyc);
  run<2>("ex2.approx.f16x2", out, cyc);
  run<3>("tanh.approx.f32", out, cyc);
  run<4>("tanh.approx.bf16x2", out, cyc);
  run<5>("rcp.approx.ftz.f32", out, cyc);
  run<6>("cvt.rn.bf16x2.f32", out, cyc);
  run<7>("ex2 + cvt.bf16x2 pairs", out, cyc);
  run<8>("int pack (2 IADD+PRMT)", out, cyc);
  run<9>("ex2 + int pack pairs", out, cyc);
  return 0;
}
