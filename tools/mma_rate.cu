// tcgen05.mma issue-rate probe: cycles per M128 x N x K16 bf16 MMA with both operands in shared memory (128B-swizzled
// K-major tiles, as the convolution kernel uses them) and with the A operand in tensor memory, N = 64 / 128 / 256, one
// CTA on every SM.  Tells how far the 64-output-channel layers can go: the tensor pipe needs N/2 cycles per MMA, the
// shared-memory operand fetch (128 x 32 B of A + N x 32 B of B) whatever the data pipe gives it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I da-clip_b200/csrc -o tools/_bin/mma_rate tools/mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace dac;

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Both operands MN-major (the P^T V context GEMM of linattn_kv.cu reads its [pixel][channel] tiles this way): M128 x N x K16
// per MMA, K = 16 pixel rows of 128 B, operands spanning two 64-channel swizzle atoms 16 KB apart.
template <int N>
__global__ void __launch_bounds__(128, 1) rate_mn_kernel(int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_tiles = smem;                       // [2 slabs][128 rows x 128 B]
  uint8_t* b_tiles = smem + 2 * 16384;           // [2 slabs][128 rows x 128 B]
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (4 * 16384) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u + i * 2654435761u % 0x00400040u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N) | (1u << 15) | (1u << 16);
    const uint64_t dk = make_sw128_desc(0);
    const uint64_t dmn = (dk & ~(static_cast<uint64_t>(0x3FFF) << 16)) | (static_cast<uint64_t>(16384 >> 4) << 16);
    const uint32_t a_lo = (smem_u32(a_tiles) & 0x3FFFF) >> 4, b_lo = (smem_u32(b_tiles) & 0x3FFFF) >> 4;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int ks = (it * 4 + k) & 7;
        umma_bf16(tmem, dmn | (a_lo + ks * 128), dmn | (b_lo + ks * 128), idesc, (it | k) ? 1u : 0u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int N>
void run_mn(int ctas) {
  const int iters = 2048, smem = 4 * 16384 + 1024;
  cudaFuncSetAttribute(rate_mn_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* d;
  cudaMalloc(&d, sizeof(long long) * ctas);
  for (int rep = 0; rep < 2; ++rep) rate_mn_kernel<N><<<ctas, 128, smem>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("MN-major N=%d: %s\n", N, cudaGetErrorString(e)); return; }
  long long* h = new long long[ctas];
  cudaMemcpy(h, d, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
  const double per = (double)mx / (4.0 * iters);
  printf("%-22s N=%3d  %7.1f cycles / MMA   tensor floor %5.1f  -> %5.1f %% of the tensor pipe,  operand fetch %4.0f B/clk\n",
         "A, B MN-major smem", N, per, N / 2.0, 100.0 * (N / 2.0) / per, (4096 + N * 32) / per);
  cudaFree(d);
  delete[] h;
}

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_tiles = smem;                       // 4 x [128 rows x 128 B]
  uint8_t* b_tile = smem + 4 * 16384;            // [N rows x 128 B]
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (4 * 16384 + N * 128) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u + i * 2654435761u % 0x00400040u;   // small finite bf16 pairs
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N);
    const uint64_t bdesc = make_sw128_desc(smem_u32(b_tile));
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint64_t adesc = make_sw128_desc(smem_u32(a_tiles + (it & 3) * 16384));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (TS) umma_bf16_ts(tmem, tmem + 256 + 8 * k + 32 * (it & 3), bdesc + 2 * k, idesc, (it | k) ? 1u : 0u);
        else umma_bf16(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) ? 1u : 0u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int N, bool TS>
void run(const char* name, int ctas) {
  const int iters = 2048, smem = 4 * 16384 + N * 128 + 1024;
  cudaFuncSetAttribute(rate_kernel<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* d;
  cudaMalloc(&d, sizeof(long long) * ctas);
  for (int rep = 0; rep < 2; ++rep) rate_kernel<N, TS><<<ctas, 128, smem>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  long long* h = new long long[ctas];
  cudaMemcpy(h, d, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
  long long mx = 0, mn = 1ll << 62;
  for (int i = 0; i < ctas; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
  const double per = (double)mx / (4.0 * iters), ideal = N / 2.0;
  printf("%-22s N=%3d  %7.1f cycles / MMA (min CTA %.1f)   tensor floor %5.1f  -> %5.1f %% of the tensor pipe,  operand fetch %4.0f B/clk\n",
         name, N, per, (double)mn / (4.0 * iters), ideal, 100.0 * ideal / per, ((TS ? 0 : 4096) + N * 32) / per);
  cudaFree(d);
  delete[] h;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d; M128 x N x K16 bf16, fp32 accumulate in TMEM, 8192 MMAs per CTA, one CTA per SM\n", sms);
  run<64, false>("A smem, B smem", sms);
  run<128, false>("A smem, B smem", sms);
  run<256, false>("A smem, B smem", sms);
  run<64, true>("A tmem, B smem", sms);
  run<128, true>("A tmem, B smem", sms);
  run<256, true>("A tmem, B smem", sms);
  run<64, false>("A smem, B smem, 1 CTA", 1);
  run_mn<128>(sms);
  run_mn<64>(sms);
  run_mn<16>(sms);
  return 0;
}
