// Softmax attention over packed qkv [B, n, 3*heads*d] bf16.
// One flash-style kernel, templated on the head dim: 16 query rows per warp, 64-key blocks of K and V staged
// row-major in shared memory with cp.async double buffering, QK^T and PV on the warp-level tensor cores (mma.sync
// m16n8k16 bf16, fp32 accumulate; B fragments via ldmatrix / ldmatrix.trans), online softmax in registers.
//  * d = 32: SpatialTransformer self-attention (attention.py:178-192), n = 1024 / 4096 tokens, 8 warps per CTA;
//    2 % of the step's FLOPs at 256^2 (the convolution GEMMs are the tcgen05 kernels).
//  * d = 64: the ViT-B/32 blocks of DA-CLIP (transformer.py:219-230), n = 50 tokens, 4 warps per (image, head);
//    with kCausal the CLIP text tower (model.py:237-249: 77 tokens under build_attention_mask's upper-triangular -inf).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <math.h>

#include "../../include/dac_b200.h"
#include "common.h"
#include "ptx.cuh"

namespace dac {

constexpr int kFaBlockK = 64;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(src_bytes)
               : "memory");
}
// 128 query rows per CTA (8 warps x 16 rows); 64-key blocks of K and V, row-major [key][d], double-buffered with
// cp.async; B fragments via ldmatrix (K) / ldmatrix.trans (V); online softmax with ex2.approx (one MUFU per score -
// at d = 32 the SFU pipe, not the tensor pipe, bounds this kernel).
template <int kFaD, int kWarps, bool kCausal = false>
__global__ void __launch_bounds__(kWarps * 32) flash_attn_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                 __nv_bfloat16* __restrict__ out, int n, int heads,
                                                                 float scale_log2) {
  constexpr int kFaBlockQ = kWarps * 16;
  constexpr int kFaPitch = kFaD + 8;   // bf16 per smem row: (2D + 16)-byte rows keep ldmatrix conflict-free
  constexpr int kKSteps = kFaD / 16;   // k-steps of QK^T
  constexpr int kNTiles = kFaD / 8;    // channel n-tiles of PV
  constexpr int kParts = kFaD / 8;     // 16-byte pieces per K / V row
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
    // 64 keys x kParts x 16 B for K and for V; rows beyond n are zero-filled
    for (int i = threadIdx.x; i < kFaBlockK * kParts * 2; i += kWarps * 32) {
      const int isv = i / (kFaBlockK * kParts), key = (i / kParts) % kFaBlockK, part = i % kParts;
      const int krow = min(k0 + key, n - 1);
      const __nv_bfloat16* src = base + static_cast<int64_t>(krow) * ld + (isv ? vcol : kcol) + part * 8;
      __nv_bfloat16* dst = (isv ? Vs[buf] : Ks[buf]) + key * kFaPitch + part * 8;
      cp_async16(dst, src, (k0 + key < n) ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  prefetch(0, 0);

  // Q fragments (A operand), kKSteps k-steps of 16 channels
  uint32_t qa[kKSteps][4];
#pragma unroll
  for (int kk = 0; kk < kKSteps; ++kk) {
    const int c = qcol + kk * 16 + cq;
    const __nv_bfloat16* p0 = base + static_cast<int64_t>(min(r0, n - 1)) * ld + c;
    const __nv_bfloat16* p1 = base + static_cast<int64_t>(min(r0 + 8, n - 1)) * ld + c;
    qa[kk][0] = *reinterpret_cast<const uint32_t*>(p0);
    qa[kk][1] = *reinterpret_cast<const uint32_t*>(p1);
    qa[kk][2] = *reinterpret_cast<const uint32_t*>(p0 + 8);
    qa[kk][3] = *reinterpret_cast<const uint32_t*>(p1 + 8);
  }
  const float sl2 = scale_log2;  // d^-0.5 * log2(e)
  float o[kNTiles][4];
#pragma unroll
  for (int j = 0; j < kNTiles; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) o[j][i] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  int buf = 0;
  // causal: key blocks past the CTA's last query row contribute nothing
  const int k_end = kCausal ? min(n, (qb + 1) * kFaBlockQ) : n;
  for (int k0 = 0; k0 < k_end; k0 += kFaBlockK, buf ^= 1) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                   // block k0 landed; everyone is done with the other buffer
    if (k0 + kFaBlockK < k_end) prefetch(buf ^ 1, k0 + kFaBlockK);
    const __nv_bfloat16* Kb = Ks[buf];
    const __nv_bfloat16* Vb = Vs[buf];

    // S = Q K^T : 8 n-tiles of 8 keys; one ldmatrix.x4 = B fragments of two k-steps
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
      for (int kp = 0; kp < kKSteps / 2; ++kp) {
        uint32_t kb[4];
        ldmatrix_x4(kb, Kb + (j * 8 + (lane & 7)) * kFaPitch + kp * 32 + (lane >> 3) * 8);
        mma_bf16_16816(s[j], qa[2 * kp], kb[0], kb[1]);
        mma_bf16_16816(s[j], qa[2 * kp + 1], kb[2], kb[3]);
      }
    }
    if (k0 + kFaBlockK > n) {  // mask keys beyond n
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int key = k0 + j * 8 + cq;
        if (key >= n) s[j][0] = s[j][2] = -INFINITY;
        if (key + 1 >= n) s[j][1] = s[j][3] = -INFINITY;
      }
    }
    if (kCausal && k0 + kFaBlockK > qb * kFaBlockQ) {  // query i sees keys <= i (key 0 is always visible)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int key = k0 + j * 8 + cq;
        if (key > r0) s[j][0] = -INFINITY;
        if (key + 1 > r0) s[j][1] = -INFINITY;
        if (key > r0 + 8) s[j][2] = -INFINITY;
        if (key + 1 > r0 + 8) s[j][3] = -INFINITY;
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
    for (int j = 0; j < kNTiles; ++j) {
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
    // O += P V : 4 k-steps of 16 keys; one ldmatrix.x4.trans = B fragments of two channel n-tiles
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int jp = 0; jp < kNTiles / 2; ++jp) {
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
  for (int jd = 0; jd < kNTiles; ++jd) {
    const int c = h * kFaD + jd * 8 + cq;
    if (r0 < n)
      *reinterpret_cast<uint32_t*>(out + (static_cast<int64_t>(b) * n + r0) * old + c) =
          pack_bf16(o[jd][0] * i0, o[jd][1] * i0);
    if (r0 + 8 < n)
      *reinterpret_cast<uint32_t*>(out + (static_cast<int64_t>(b) * n + r0 + 8) * old + c) =
          pack_bf16(o[jd][2] * i1, o[jd][3] * i1);
  }
}

}  // namespace dac

using namespace dac;

int dac_attention_tc(const void* qkv, void* out, int B, int n, int heads, cudaStream_t stream);   // attention_tc.cu
int dac_attention_tc2(const void* qkv, void* out, int B, int n, int heads, cudaStream_t stream);  // attention_tc2.cu
int dac_attention_vit(const void* qkv, void* out, int B, int n, int heads, int causal, cudaStream_t stream);  // attention_vit.cu

extern "C" int dac_attention(const void* qkv, void* out, int32_t B, int32_t n, int32_t heads, int32_t d,
                             dac_stream_t stream) {
  if (!qkv || !out) return set_error(-1, "dac_attention: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // d = 32 with whole 128-token tiles and head pairs: the tcgen05 kernel; everything else: the mma.sync kernel below
  static const bool no_tc = getenv("DAC_NO_TC_ATTN") != nullptr;
  if (d == 32 && n % 128 == 0 && heads % 2 == 0 && B > 0 && !no_tc) {
    // DAC_ATTN_V1=1: round 1's kernel (one softmax warp per head and lane quadrant), kept for A/B timing
    if (getenv("DAC_ATTN_V1")) return dac_attention_tc(qkv, out, B, n, heads, s);
    return dac_attention_tc2(qkv, out, B, n, heads, s);
  }
  const float l2e = 1.4426950408889634f;
  if (d == 32) {   // UNet self-attention: 1024 / 4096 tokens, 128 query rows per CTA
    flash_attn_kernel<32, 8><<<dim3((n + 127) / 128, heads, B), 256, 0, s>>>(
        static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), n, heads, 0.17677669529663687f * l2e);
    return check_launch("flash_attn_kernel<32>");
  }
  // d = 64, up to 512 tokens (the ViT blocks: 50 / 257 tokens): tcgen05 kernel of attention_vit.cu
  if (d == 64 && n > 0 && n <= 512 && B > 0 && heads > 0 && !no_tc) return dac_attention_vit(qkv, out, B, n, heads, 0, s);
  if (d == 64) {   // ViT-B/32 blocks: 50 tokens, one 64-row CTA per (image, head)
    flash_attn_kernel<64, 4><<<dim3((n + 63) / 64, heads, B), 128, 0, s>>>(
        static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), n, heads, 0.125f * l2e);
    return check_launch("flash_attn_kernel<64>");
  }
  return set_error(-2, "dac_attention: unsupported head dim %d / length %d", d, n);
}

extern "C" int dac_attention_causal(const void* qkv, void* out, int32_t B, int32_t n, int32_t heads, int32_t d,
                                    dac_stream_t stream) {
  if (!qkv || !out) return set_error(-1, "dac_attention_causal: null argument");
  if (d != 64 || B <= 0 || n <= 0 || heads <= 0)
    return set_error(-2, "dac_attention_causal: head dim 64 only (got %d), positive sizes", d);
  // one key block (the text tower: 77 tokens): tcgen05 kernel of attention_vit.cu with the causal mask
  if (n <= 128 && !getenv("DAC_NO_TC_ATTN"))
    return dac_attention_vit(qkv, out, B, n, heads, 1, static_cast<cudaStream_t>(stream));
  flash_attn_kernel<64, 4, true><<<dim3((n + 63) / 64, heads, B), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), n, heads, 0.125f * 1.4426950408889634f);
  return check_launch("flash_attn_kernel<64, causal>");
}
