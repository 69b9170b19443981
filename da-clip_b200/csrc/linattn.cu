// LinearAttention (module_util.py:157-185) context reduction.
//   ctx[b,h,d,e] = sum_n softmax_n(k[b,h,d,:])[n] * v[b,h,e,n] / hw
// Pass 1 (linattn_context_kernel): each CTA walks a slab of pixels of one (image, head) with an online
// (running max / rescale) softmax over the pixel axis and writes a partial {C[32][32], m[32], S[32]}.
// Pass 2 (linattn_fold_kernel): merges the partials and folds ctx into the to_out 1x1 weight so that the
// "apply context + to_out" pair becomes ONE tensor-core GEMM with a per-image weight:
//   weff[b][c][h*32+d] = sum_e Wout[c][h*32+e] * ctx[b,h,d,e].
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/dac_b200.h"
#include "common.h"
#include "ptx.cuh"

namespace dac {

constexpr int kLaP = 64;                 // pixels per sub-tile
constexpr int kLaPitch = kLaP + 8;       // bf16 pitch of the transposed [channel][pixel] tiles
constexpr int kPartial = 32 * 32 + 64;

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// One CTA (4 warps) per (pixel slab, head, image).  Per 64-pixel sub-tile: k and v are scattered transposed
// ([channel][pixel], bf16) into shared memory; the running column max / rescale / exp / row sums are done by
// (channel, quarter) threads on 16 contiguous pixels; C[d][e] += sum_p exp(k)[d][p] v[e][p] runs on the warp-level
// tensor cores (mma.sync m16n8k16 bf16, fp32 accumulate): warp w owns d rows 16*(w/2).. and e cols 16*(w%2)...
__global__ void __launch_bounds__(128) linattn_context_kernel(const __nv_bfloat16* __restrict__ qkv, int hw,
                                                              int slab, float* __restrict__ partial) {
  __shared__ __align__(16) __nv_bfloat16 Kt[32 * kLaPitch];
  __shared__ __align__(16) __nv_bfloat16 Vt[32 * kLaPitch];
  __shared__ float m_run[32], s_run[32], scale[32];

  const int chunk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nchunks = gridDim.x;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int p_begin = chunk * slab, p_end = min(hw, p_begin + slab);
  const __nv_bfloat16* base = qkv + static_cast<int64_t>(b) * hw * 384;
  const int d0 = 16 * (warp >> 1), e0 = 16 * (warp & 1);
  float acc[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  if (t < 32) {
    m_run[t] = -INFINITY;
    s_run[t] = 0.f;
  }
  const __nv_bfloat16 neg_inf = __float2bfloat16(-INFINITY);

  for (int p0 = p_begin; p0 < p_end; p0 += kLaP) {
    __syncthreads();  // previous sub-tile fully consumed (and m_run / s_run initialised)
    {  // load 64 pixels x (32 k + 32 v) channels; thread -> (pixel, 16-channel half); scatter transposed
      const int pl = t >> 1, part = t & 1;
      const int p = p0 + pl;
      uint4 uk[2], uv[2];
      if (p < p_end) {
        const __nv_bfloat16* row = base + static_cast<int64_t>(p) * 384;
        uk[0] = __ldg(reinterpret_cast<const uint4*>(row + 128 + h * 32 + part * 16));
        uk[1] = __ldg(reinterpret_cast<const uint4*>(row + 128 + h * 32 + part * 16) + 1);
        uv[0] = __ldg(reinterpret_cast<const uint4*>(row + 256 + h * 32 + part * 16));
        uv[1] = __ldg(reinterpret_cast<const uint4*>(row + 256 + h * 32 + part * 16) + 1);
      }
      const __nv_bfloat16* kk = reinterpret_cast<const __nv_bfloat16*>(uk);
      const __nv_bfloat16* vv = reinterpret_cast<const __nv_bfloat16*>(uv);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int c = part * 16 + j;
        Kt[c * kLaPitch + pl] = (p < p_end) ? kk[j] : neg_inf;           // masked pixel: exp -> 0
        Vt[c * kLaPitch + pl] = (p < p_end) ? vv[j] : __float2bfloat16(0.f);
      }
    }
    __syncthreads();
    {  // thread -> (channel d = t/4, 16-pixel quarter): running max, rescale factor, exp in place, row sum
      const int d = t >> 2, qtr = t & 3;
      __nv_bfloat16* rowp = &Kt[d * kLaPitch + qtr * 16];
      float kv[16];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint4 u = *reinterpret_cast<const uint4*>(rowp + q * 8);
        float2 f;
        f = unpack_bf16(u.x); kv[q * 8 + 0] = f.x; kv[q * 8 + 1] = f.y;
        f = unpack_bf16(u.y); kv[q * 8 + 2] = f.x; kv[q * 8 + 3] = f.y;
        f = unpack_bf16(u.z); kv[q * 8 + 4] = f.x; kv[q * 8 + 5] = f.y;
        f = unpack_bf16(u.w); kv[q * 8 + 6] = f.x; kv[q * 8 + 7] = f.y;
      }
      float m = kv[0];
#pragma unroll
      for (int j = 1; j < 16; ++j) m = fmaxf(m, kv[j]);
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      const float mo = m_run[d];
      const float mn = fmaxf(mo, m);
      // first sub-tile of a slab always holds at least one live pixel, so mn is finite
      float s = 0.f;
      uint32_t packed[8];
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const float ea = __expf(kv[j] - mn), eb = __expf(kv[j + 1] - mn);
        packed[j >> 1] = pack_bf16(ea, eb);
        const float2 r = unpack_bf16(packed[j >> 1]);      // sum exactly what the tensor cores will multiply
        s += r.x + r.y;
      }
      *reinterpret_cast<uint4*>(rowp) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
      *reinterpret_cast<uint4*>(rowp + 8) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      __syncwarp();
      if (qtr == 0) {
        const float sc = (mo == -INFINITY) ? 0.f : __expf(mo - mn);
        scale[d] = sc;
        m_run[d] = mn;
        s_run[d] = s_run[d] * sc + s;
      }
    }
    __syncthreads();
    {  // C[d][e] = C[d][e] * scale[d] + sum_p P[d][p] V[e][p]
      const int r = lane >> 2, cq = 2 * (lane & 3);
      const float sc0 = scale[d0 + r], sc1 = scale[d0 + r + 8];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        acc[i][0] *= sc0; acc[i][1] *= sc0; acc[i][2] *= sc1; acc[i][3] *= sc1;
      }
#pragma unroll
      for (int ks = 0; ks < kLaP / 16; ++ks) {
        uint32_t a[4];
        const __nv_bfloat16* ap = &Kt[(d0 + r) * kLaPitch + ks * 16 + cq];
        a[0] = *reinterpret_cast<const uint32_t*>(ap);
        a[1] = *reinterpret_cast<const uint32_t*>(ap + 8 * kLaPitch);
        a[2] = *reinterpret_cast<const uint32_t*>(ap + 8);
        a[3] = *reinterpret_cast<const uint32_t*>(ap + 8 * kLaPitch + 8);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const __nv_bfloat16* bp = &Vt[(e0 + i * 8 + r) * kLaPitch + ks * 16 + cq];
          mma_bf16_16816(acc[i], a, *reinterpret_cast<const uint32_t*>(bp), *reinterpret_cast<const uint32_t*>(bp + 8));
        }
      }
    }
  }
  __syncthreads();
  float* out = partial + ((static_cast<int64_t>(b) * 4 + h) * nchunks + chunk) * kPartial;
  {
    const int r = lane >> 2, cq = 2 * (lane & 3);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int e = e0 + i * 8 + cq;
      out[(d0 + r) * 32 + e] = acc[i][0];
      out[(d0 + r) * 32 + e + 1] = acc[i][1];
      out[(d0 + r + 8) * 32 + e] = acc[i][2];
      out[(d0 + r + 8) * 32 + e + 1] = acc[i][3];
    }
  }
  if (t < 32) {
    out[1024 + t] = m_run[t];
    out[1056 + t] = s_run[t];
  }
}

__global__ void __launch_bounds__(256) linattn_fold_kernel(const float* __restrict__ partial, int hw, int nchunks,
                                                           const float* __restrict__ w_out, int C, int c_pad,
                                                           __nv_bfloat16* __restrict__ weff) {
  __shared__ float ctx[32 * 33];
  __shared__ float wgt[128 * 32];
  __shared__ float inv_s[32];
  const int h = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
  const float* pbase = partial + (static_cast<int64_t>(b) * 4 + h) * nchunks * kPartial;
  if (t < 32) {
    float M = -INFINITY;
    for (int c = 0; c < nchunks; ++c) M = fmaxf(M, pbase[c * kPartial + 1024 + t]);
    float S = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      const float mc = pbase[c * kPartial + 1024 + t];
      const float w = (mc == -INFINITY) ? 0.f : __expf(mc - M);
      wgt[c * 32 + t] = w;
      S += pbase[c * kPartial + 1056 + t] * w;
    }
    inv_s[t] = 1.0f / (S * static_cast<float>(hw));   // softmax denominator and v / (h*w) (module_util.py:177)
  }
  __syncthreads();
  for (int i = t; i < 1024; i += 256) {
    const int d = i >> 5, e = i & 31;
    float a = 0.f;
    for (int c = 0; c < nchunks; ++c) a += pbase[c * kPartial + i] * wgt[c * 32 + d];
    ctx[d * 33 + e] = a * inv_s[d];
  }
  __syncthreads();
  const int d = t & 31;
  for (int c = t >> 5; c < C; c += 8) {
    const float* wr = w_out + static_cast<int64_t>(c) * 128 + h * 32;
    float a = 0.f;
#pragma unroll
    for (int e = 0; e < 32; ++e) a += __ldg(wr + e) * ctx[d * 33 + e];
    weff[(static_cast<int64_t>(b) * c_pad + c) * 128 + h * 32 + d] = __float2bfloat16(a);
  }
}

}  // namespace dac

using namespace dac;

extern "C" int dac_linattn_context(const void* qkv, int32_t B, int32_t hw, int32_t nchunks, float* partial,
                                   dac_stream_t stream) {
  if (!qkv || !partial) return set_error(-1, "dac_linattn_context: null argument");
  if (nchunks < 1 || nchunks > 128) return set_error(-2, "dac_linattn_context: nchunks must be in [1,128]");
  const int slab = static_cast<int>(ceil_div(ceil_div(hw, nchunks), kLaP) * kLaP);
  linattn_context_kernel<<<dim3(nchunks, 4, B), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(qkv), hw, slab, partial);
  return check_launch("linattn_context_kernel");
}

extern "C" int dac_linattn_fold(const float* partial, int32_t B, int32_t hw, int32_t nchunks, const float* w_out,
                                int32_t C, int32_t c_pad, void* weff, dac_stream_t stream) {
  if (!partial || !w_out || !weff) return set_error(-1, "dac_linattn_fold: null argument");
  if (nchunks < 1 || nchunks > 128) return set_error(-2, "dac_linattn_fold: nchunks must be in [1,128]");
  linattn_fold_kernel<<<dim3(4, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      partial, hw, nchunks, w_out, C, c_pad, static_cast<__nv_bfloat16*>(weff));
  return check_launch("linattn_fold_kernel");
}
