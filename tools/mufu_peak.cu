// Measures the ex2.approx throughput of one SM-filling launch (ops / clk / SM).  nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters, long long* cycles) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a[i] * 0.5f - 1.0f;
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMallocManaged(&cyc, 8);
  for (int threads : {128, 256, 512, 1024}) {
    const int iters = 4096;
    k<<<148, threads>>>(out, iters, cyc); cudaDeviceSynchronize();
    k<<<148, threads>>>(out, iters, cyc); cudaDeviceSynchronize();
    printf("threads/SM %4d: %.2f ex2/clk/SM\n", threads, (double)threads * iters * 8 / (double)*cyc);
  }
  return 0;
}
