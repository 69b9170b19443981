// Pieces shared by the tcgen05 kernels of this library (conv_kernel.cuh, linattn_qout.cu): tile / role constants, the
// contiguous tile range of a persistent CTA, and the 32-column epilogue chunk helpers (TMEM -> registers -> global or
// the 128B-swizzled staging tile).
#pragma once
#include <cuda_runtime.h>

#include "ptx.cuh"

namespace dac {

constexpr int kTileM = 128;
constexpr int kChunkK = 64;   // bf16 per K step = one 128 B swizzle row
constexpr int kThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int kEpiWarps = 8;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccStride = 256;
constexpr uint32_t kABytes = kTileM * kChunkK * 2;  // 16 KB
constexpr int kMaxStages = 8;

// CTA b owns the contiguous tile range [total*b/grid, total*(b+1)/grid): consecutive tiles of a CTA belong to the
// same image (FiLM parameters stay cached) and neighbouring rows (halo re-reads hit L2).
__device__ __forceinline__ void tile_range(int total, int& begin, int& end) {
  begin = static_cast<int>(static_cast<long long>(total) * blockIdx.x / gridDim.x);
  end = static_cast<int>(static_cast<long long>(total) * (blockIdx.x + 1) / gridDim.x);
}

// Inverse of tile_range: the CTA whose range holds tile t (ceil((t + 1) * grid / total) - 1).
__host__ __device__ __forceinline__ int tile_owner(int t, int total, int grid) {
  return static_cast<int>((static_cast<long long>(t + 1) * grid + total - 1) / total) - 1;
}
// Most CTAs any image spans when every image is tpi consecutive tiles of B * tpi (see dac_linattn_ctx_slots).
inline int max_image_span(int B, int tpi, int grid) {
  int span = 1;
  for (int b = 0; b < B; ++b) {
    const int s = tile_owner((b + 1) * tpi - 1, B * tpi, grid) - tile_owner(b * tpi, B * tpi, grid) + 1;
    span = s > span ? s : span;
  }
  return span;
}

// ---- 32-column chunk helpers: everything statically indexed so the chunk lives in registers ----
__device__ __forceinline__ void chunk_from_tmem(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  tmem_ld32(taddr, r);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void chunk_add_f32(const float* __restrict__ src, float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + q);
    v[4 * q] += a.x; v[4 * q + 1] += a.y; v[4 * q + 2] += a.z; v[4 * q + 3] += a.w;
  }
}
__device__ __forceinline__ void chunk_film(const float* __restrict__ sc, const float* __restrict__ sh,
                                           float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(sc) + q);
    const float4 b = __ldg(reinterpret_cast<const float4*>(sh) + q);
    v[4 * q] = fmaf(v[4 * q], a.x + 1.0f, b.x);
    v[4 * q + 1] = fmaf(v[4 * q + 1], a.y + 1.0f, b.y);
    v[4 * q + 2] = fmaf(v[4 * q + 2], a.z + 1.0f, b.z);
    v[4 * q + 3] = fmaf(v[4 * q + 3], a.w + 1.0f, b.w);
  }
}
__device__ __forceinline__ void chunk_add_bf16(const __nv_bfloat16* __restrict__ src, float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {          // 2 x 32 B = this row's 32 channels
    const U32x8 u = ldg256(src + q * 16);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 a = unpack_bf16(u.v[j]);
      v[q * 16 + 2 * j] += a.x;
      v[q * 16 + 2 * j + 1] += a.y;
    }
  }
}
__device__ __forceinline__ void chunk_store_bf16(__nv_bfloat16* __restrict__ dst, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    U32x8 u;
#pragma unroll
    for (int j = 0; j < 8; ++j) u.v[j] = pack_bf16(v[q * 16 + 2 * j], v[q * 16 + 2 * j + 1]);
    stg256(dst + q * 16, u);
  }
}
// 32 columns of one row into the staging tile: 64-channel slabs of 128 rows x 128 B, 16 B pieces XOR-swizzled by the
// row (the SWIZZLE_128B pattern of the output tensor map; also conflict-free for one-row-per-lane writes).
__device__ __forceinline__ void chunk_stage_bf16(uint8_t* stg, int row, int col, const float (&v)[32]) {
  // shared-state-space stores (STS.128 with a 32-bit address): a generic ST.E.128 costs a 64-bit address add per store
  const uint32_t slab = smem_u32(stg) + (col >> 6) * (kTileM * 128) + row * 128;
  const int c16 = (col & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slab + (((c16 + q) ^ (row & 7)) << 4)),
                 "r"(pack_bf16(v[q * 8 + 0], v[q * 8 + 1])), "r"(pack_bf16(v[q * 8 + 2], v[q * 8 + 3])),
                 "r"(pack_bf16(v[q * 8 + 4], v[q * 8 + 5])), "r"(pack_bf16(v[q * 8 + 6], v[q * 8 + 7]))
                 : "memory");
  }
}
// Same, addressed in the shared state space (32-bit address, STS.128): the generic-pointer form above costs a 64-bit add
// per store, which matters in the issue-bound softmax loops.
__device__ __forceinline__ void chunk_stage_bf16_s(uint32_t stg_s, int row, int col, const float (&v)[32]) {
  const uint32_t slab = stg_s + (col >> 6) * (kTileM * 128) + row * 128;
  const int c16 = (col & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slab + (((c16 + q) ^ (row & 7)) << 4)),
                 "r"(pack_bf16(v[q * 8 + 0], v[q * 8 + 1])), "r"(pack_bf16(v[q * 8 + 2], v[q * 8 + 3])),
                 "r"(pack_bf16(v[q * 8 + 4], v[q * 8 + 5])), "r"(pack_bf16(v[q * 8 + 6], v[q * 8 + 7]))
                 : "memory");
  }
}
// ... and the read side: add the 32 staged bf16 values of this row (a TMA-loaded residual tile) to v
__device__ __forceinline__ void chunk_add_staged(const uint8_t* stg, int row, int col, float (&v)[32]) {
  const uint32_t slab = smem_u32(stg) + (col >> 6) * (kTileM * 128) + row * 128;
  const int c16 = (col & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t ux, uy, uz, uw;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(ux), "=r"(uy), "=r"(uz), "=r"(uw)
                 : "r"(slab + (((c16 + q) ^ (row & 7)) << 4)));
    const float2 a = unpack_bf16(ux), b = unpack_bf16(uy), c = unpack_bf16(uz), d = unpack_bf16(uw);
    unpack_f32x2(add_f32x2(pack_f32x2(v[q * 8 + 0], v[q * 8 + 1]), pack_f32x2(a.x, a.y)), v[q * 8 + 0], v[q * 8 + 1]);
    unpack_f32x2(add_f32x2(pack_f32x2(v[q * 8 + 2], v[q * 8 + 3]), pack_f32x2(b.x, b.y)), v[q * 8 + 2], v[q * 8 + 3]);
    unpack_f32x2(add_f32x2(pack_f32x2(v[q * 8 + 4], v[q * 8 + 5]), pack_f32x2(c.x, c.y)), v[q * 8 + 4], v[q * 8 + 5]);
    unpack_f32x2(add_f32x2(pack_f32x2(v[q * 8 + 6], v[q * 8 + 7]), pack_f32x2(d.x, d.y)), v[q * 8 + 6], v[q * 8 + 7]);
  }
}
__device__ __forceinline__ void chunk_store_f32(float* __restrict__ dst, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q)
    reinterpret_cast<float4*>(dst)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

}  // namespace dac
