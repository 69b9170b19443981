// Softmax attention over packed qkv [B, n, 3*heads*d] bf16.
//  * d = 32 (SpatialTransformer self-attention, attention.py:178-192; n = 1024 / 4096): flash-style kernel,
//    64 query rows per CTA (16 per warp), 64-key blocks staged in shared memory, QK^T and PV on the
//    warp-level tensor-core path (mma.sync m16n8k16 bf16, fp32 accumulate), online softmax in registers.
//    It is 2 % of the step's FLOPs at 256^2; the convolution GEMMs are the tcgen05 kernels.
//  * d = 64, n <= 64 (ViT-B/32 blocks of DA-CLIP, transformer.py:219-230; n = 50): one CTA per (image, head),
//    fp32 CUDA-core math (the whole K/V of a head is 25 KB).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/dac_b200.h"
#include "common.h"
#include "ptx.cuh"

namespace dac {

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kFaBlockQ = 128, kFaBlockK = 64, kFaD = 32;
constexpr int kFaPitch = kFaD + 8;  // bf16 elements per smem row: 80 B rows keep ldmatrix conflict-free

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem)));
}

// 128 query rows per CTA (8 warps x 16 rows); 64-key blocks of K and V, row-major [key][d], double-buffered with
// cp.async; B fragments via ldmatrix (K) / ldmatrix.trans (V); online softmax with ex2.approx (one MUFU per score -
// at d = 32 the SFU pipe, not the tensor pipe, bounds this kernel).
__global__ void __launch_bounds__(256) flash_attn_d32_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                             __nv_bfloat16* __restrict__ out, int n, int heads) {
  __shared__ __align__(16) __nv_bfloat16 Ks[2][kFaBlockK * kFaPitch];
  __shared__ __align__(16) __nv_bfloat16 Vs[2][kFaBlockK * kFaPitch];
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = 3 * heads * kFaD;
  const __nv_bfloat16* base = qkv + static_cast<int64_t>(b) * n * ld;
  const int qcol = h * kFaD, kcol = heads * kFaD + h * kFaD, vcol = 2 * heads * kFaD + h * kFaD;
  const int r0 = qb * kFaBlockQ + warp * 16 + (lane >> 2);  // rows r0 and r0 + 8
  const int cq = 2 * (lane & 3);

  auto prefetch = [&](int buf, int k0) {
    // 64 keys x 4 x 16 B for K and for V: 512 copies, two per thread; rows beyond n are zero-filled
    for (int i = threadIdx.x; i < kFaBlockK * 8; i += 256) {
      const int isv = i >> 8, key = (i >> 2) & 63, part = i & 3;
      const int krow = min(k0 + key, n - 1);
      const __nv_bfloat16* src = base + static_cast<int64_t>(krow) * ld + (isv ? vcol : kcol) + part * 8;
      __nv_bfloat16* dst = (isv ? Vs[buf] : Ks[buf]) + key * kFaPitch + part * 8;
      cp_async16(dst, src, (k0 + key < n) ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  prefetch(0, 0);

  // Q fragments (A operand), 2 k-steps of 16 channels
  uint32_t qa[2][4];
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    const int c = qcol + kk * 16 + cq;
    const __nv_bfloat16* p0 = base + static_cast<int64_t>(min(r0, n - 1)) * ld + c;
    const __nv_bfloat16* p1 = base + static_cast<int64_t>(min(r0 + 8, n - 1)) * ld + c;
    qa[kk][0] = *reinterpret_cast<const uint32_t*>(p0);
    qa[kk][1] = *reinterpret_cast<const uint32_t*>(p1);
    qa[kk][2] = *reinterpret_cast<const uint32_t*>(p0 + 8);
    qa[kk][3] = *reinterpret_cast<const uint32_t*>(p1 + 8);
  }
  const float sl2 = 0.17677669529663687f * 1.4426950408889634f;  // d^-0.5 * log2(e)
  float o[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) o[j][i] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  int buf = 0;
  for (int k0 = 0; k0 < n; k0 += kFaBlockK, buf ^= 1) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                   // block k0 landed; everyone is done with the other buffer
    if (k0 + kFaBlockK < n) prefetch(buf ^ 1, k0 + kFaBlockK);
    const __nv_bfloat16* Kb = Ks[buf];
    const __nv_bfloat16* Vb = Vs[buf];

    // S = Q K^T : 8 n-tiles of 8 keys; one ldmatrix.x4 per n-tile = B fragments of both k-steps
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t kb[4];
      ldmatrix_x4(kb, Kb + (j * 8 + (lane & 7)) * kFaPitch + (lane >> 3) * 8);
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      mma_bf16_16816(s[j], qa[0], kb[0], kb[1]);
      mma_bf16_16816(s[j], qa[1], kb[2], kb[3]);
    }
    if (k0 + kFaBlockK > n) {  // mask keys beyond n
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int key = k0 + j * 8 + cq;
        if (key >= n) s[j][0] = s[j][2] = -INFINITY;
        if (key + 1 >= n) s[j][1] = s[j][3] = -INFINITY;
      }
    }
    // online softmax (rows r0: regs 0,1; r0+8: regs 2,3)
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float a0 = ex2_approx((m0 - mx0) * sl2), a1 = ex2_approx((m1 - mx1) * sl2);  // -inf on block 0 -> 0
    m0 = mx0;
    m1 = mx1;
    const float mb0 = mx0 * sl2, mb1 = mx1 * sl2;
    l0 *= a0;
    l1 *= a1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o[j][0] *= a0; o[j][1] *= a0; o[j][2] *= a1; o[j][3] *= a1;
    }
    uint32_t pa[4][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = ex2_approx(fmaf(s[j][0], sl2, -mb0)), p1 = ex2_approx(fmaf(s[j][1], sl2, -mb0));
      const float p2 = ex2_approx(fmaf(s[j][2], sl2, -mb1)), p3 = ex2_approx(fmaf(s[j][3], sl2, -mb1));
      l0 += p0 + p1;
      l1 += p2 + p3;
      pa[j >> 1][(j & 1) * 2 + 0] = pack_bf16(p0, p1);
      pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
    // O += P V : 4 k-steps of 16 keys; per k-step two ldmatrix.x4.trans = B fragments of the 4 channel n-tiles
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int jp = 0; jp < 2; ++jp) {
        uint32_t vb[4];
        ldmatrix_x4_trans(vb, Vb + (ks * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * kFaPitch +
                                  (jp * 2 + (lane >> 4)) * 8);
        mma_bf16_16816(o[jp * 2], pa[ks], vb[0], vb[1]);
        mma_bf16_16816(o[jp * 2 + 1], pa[ks], vb[2], vb[3]);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int old = heads * kFaD;
#pragma unroll
  for (int jd = 0; jd < 4; ++jd) {
    const int c = h * kFaD + jd * 8 + cq;
    if (r0 < n)
      *reinterpret_cast<uint32_t*>(out + (static_cast<int64_t>(b) * n + r0) * old + c) =
          pack_bf16(o[jd][0] * i0, o[jd][1] * i0);
    if (r0 + 8 < n)
      *reinterpret_cast<uint32_t*>(out + (static_cast<int64_t>(b) * n + r0 + 8) * old + c) =
          pack_bf16(o[jd][2] * i1, o[jd][3] * i1);
  }
}

// ------------------------------------------------------------------------------------------------ ViT (d=64, n<=64)
__global__ void __launch_bounds__(128) small_attn_d64_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                             __nv_bfloat16* __restrict__ out, int n, int heads) {
  constexpr int D = 64, P = D + 1;
  __shared__ float Ksh[64 * P], Vsh[64 * P], Qsh[4][D], Psh[4][64];
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = 3 * heads * D;
  const __nv_bfloat16* base = qkv + static_cast<int64_t>(b) * n * ld;
  for (int i = threadIdx.x; i < n * D; i += 128) {
    const int key = i / D, c = i % D;
    Ksh[key * P + c] = __bfloat162float(base[static_cast<int64_t>(key) * ld + heads * D + h * D + c]);
    Vsh[key * P + c] = __bfloat162float(base[static_cast<int64_t>(key) * ld + 2 * heads * D + h * D + c]);
  }
  __syncthreads();
  for (int q = warp; q < n; q += 4) {
    Qsh[warp][lane] = __bfloat162float(base[static_cast<int64_t>(q) * ld + h * D + lane]);
    Qsh[warp][lane + 32] = __bfloat162float(base[static_cast<int64_t>(q) * ld + h * D + lane + 32]);
    __syncwarp();
    float s0 = -INFINITY, s1 = -INFINITY;
    if (lane < n) {
      float a = 0.f;
#pragma unroll 8
      for (int c = 0; c < D; ++c) a += Qsh[warp][c] * Ksh[lane * P + c];
      s0 = a * 0.125f;
    }
    if (lane + 32 < n) {
      float a = 0.f;
#pragma unroll 8
      for (int c = 0; c < D; ++c) a += Qsh[warp][c] * Ksh[(lane + 32) * P + c];
      s1 = a * 0.125f;
    }
    float m = fmaxf(s0, s1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float p0 = (lane < n) ? expf(s0 - m) : 0.f, p1 = (lane + 32 < n) ? expf(s1 - m) : 0.f;
    float l = p0 + p1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    Psh[warp][lane] = p0;
    Psh[warp][lane + 32] = p1;
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int key = 0; key < n; ++key) {
      const float p = Psh[warp][key];
      o0 += p * Vsh[key * P + lane];
      o1 += p * Vsh[key * P + lane + 32];
    }
    const float inv = 1.0f / l;
    __nv_bfloat16* dst = out + (static_cast<int64_t>(b) * n + q) * (heads * D) + h * D;
    dst[lane] = __float2bfloat16(o0 * inv);
    dst[lane + 32] = __float2bfloat16(o1 * inv);
    __syncwarp();
  }
}

}  // namespace dac

using namespace dac;

extern "C" int dac_attention(const void* qkv, void* out, int32_t B, int32_t n, int32_t heads, int32_t d,
                             dac_stream_t stream) {
  if (!qkv || !out) return set_error(-1, "dac_attention: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (d == 32) {
    flash_attn_d32_kernel<<<dim3((n + kFaBlockQ - 1) / kFaBlockQ, heads, B), 256, 0, s>>>(
        static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), n, heads);
    return check_launch("flash_attn_d32_kernel");
  }
  if (d == 64 && n <= 64) {
    small_attn_d64_kernel<<<dim3(heads, B), 128, 0, s>>>(static_cast<const __nv_bfloat16*>(qkv),
                                                         static_cast<__nv_bfloat16*>(out), n, heads);
    return check_launch("small_attn_d64_kernel");
  }
  return set_error(-2, "dac_attention: unsupported head dim %d / length %d", d, n);
}
