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

constexpr int kLaP = 64;       // pixels per sub-tile
constexpr int kLaKPitch = 34;  // fp32 pitch of the exp(k) tile (even: float2 loads)
constexpr int kPartial = 32 * 32 + 64;

__global__ void __launch_bounds__(256) linattn_context_kernel(const __nv_bfloat16* __restrict__ qkv, int hw,
                                                              int slab, float* __restrict__ partial) {
  __shared__ __align__(16) float ks[kLaP * kLaKPitch];
  __shared__ __align__(16) float vs[kLaP * 32];
  __shared__ float red[8 * 32];
  __shared__ float m_run[32], s_run[32], scale[32];
  __shared__ float cred[32 * 32];

  const int chunk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nchunks = gridDim.x;
  const int t = threadIdx.x;
  const int p_begin = chunk * slab, p_end = min(hw, p_begin + slab);
  const __nv_bfloat16* base = qkv + static_cast<int64_t>(b) * hw * 384;

  // compute-phase mapping: 4 pixel groups x (16 d-pairs x 4 e-octets)
  const int pg4 = t >> 6, u = t & 63, dp = u >> 2, eo = u & 3;
  const int d0 = 2 * dp;
  float acc[2][8];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  if (t < 32) {
    m_run[t] = -INFINITY;
    s_run[t] = 0.f;
  }
  __syncthreads();

  for (int p0 = p_begin; p0 < p_end; p0 += kLaP) {
    {  // load 64 pixels x (32 k + 32 v) channels: thread -> (pixel, 8-channel part)
      const int pl = t >> 2, part = t & 3;
      const int p = p0 + pl;
      float kv[8], vv[8];
      if (p < p_end) {
        const __nv_bfloat16* row = base + static_cast<int64_t>(p) * 384;
        const uint4 uk = *reinterpret_cast<const uint4*>(row + 128 + h * 32 + part * 8);
        const uint4 uv = *reinterpret_cast<const uint4*>(row + 256 + h * 32 + part * 8);
        float2 f;
        f = unpack_bf16(uk.x); kv[0] = f.x; kv[1] = f.y;
        f = unpack_bf16(uk.y); kv[2] = f.x; kv[3] = f.y;
        f = unpack_bf16(uk.z); kv[4] = f.x; kv[5] = f.y;
        f = unpack_bf16(uk.w); kv[6] = f.x; kv[7] = f.y;
        f = unpack_bf16(uv.x); vv[0] = f.x; vv[1] = f.y;
        f = unpack_bf16(uv.y); vv[2] = f.x; vv[3] = f.y;
        f = unpack_bf16(uv.z); vv[4] = f.x; vv[5] = f.y;
        f = unpack_bf16(uv.w); vv[6] = f.x; vv[7] = f.y;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { kv[j] = -INFINITY; vv[j] = 0.f; }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ks[pl * kLaKPitch + part * 8 + j] = kv[j];
        vs[pl * 32 + part * 8 + j] = vv[j];
      }
    }
    __syncthreads();
    {  // tile max per d: thread -> (d = t%32, 8-pixel group t/32)
      const int d = t & 31, g = t >> 5;
      float m = -INFINITY;
#pragma unroll
      for (int i = 0; i < 8; ++i) m = fmaxf(m, ks[(g * 8 + i) * kLaKPitch + d]);
      red[g * 32 + d] = m;
    }
    __syncthreads();
    if (t < 32) {
      float m = red[t];
#pragma unroll
      for (int g = 1; g < 8; ++g) m = fmaxf(m, red[g * 32 + t]);
      const float mo = m_run[t];
      const float mn = fmaxf(mo, m);
      scale[t] = (mo == -INFINITY) ? 0.f : __expf(mo - mn);
      m_run[t] = mn;
    }
    __syncthreads();
    {  // exponentiate in place + partial row sums
      const int d = t & 31, g = t >> 5;
      const float mn = m_run[d];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = (g * 8 + i) * kLaKPitch + d;
        const float e = __expf(ks[idx] - mn);  // -inf (masked pixel) -> 0
        ks[idx] = e;
        s += e;
      }
      red[g * 32 + d] = s;
    }
    __syncthreads();
    if (t < 32) {
      float s = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) s += red[g * 32 + t];
      s_run[t] = s_run[t] * scale[t] + s;
    }
    {  // C[d][e] += sum_p ek[p][d] v[p][e]
      const float sc0 = scale[d0], sc1 = scale[d0 + 1];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[0][j] *= sc0;
        acc[1][j] *= sc1;
      }
      for (int p = pg4; p < kLaP; p += 4) {
        const float2 kk = *reinterpret_cast<const float2*>(&ks[p * kLaKPitch + d0]);
        const float4 va = *reinterpret_cast<const float4*>(&vs[p * 32 + eo * 8]);
        const float4 vb = *reinterpret_cast<const float4*>(&vs[p * 32 + eo * 8 + 4]);
        acc[0][0] += kk.x * va.x; acc[0][1] += kk.x * va.y; acc[0][2] += kk.x * va.z; acc[0][3] += kk.x * va.w;
        acc[0][4] += kk.x * vb.x; acc[0][5] += kk.x * vb.y; acc[0][6] += kk.x * vb.z; acc[0][7] += kk.x * vb.w;
        acc[1][0] += kk.y * va.x; acc[1][1] += kk.y * va.y; acc[1][2] += kk.y * va.z; acc[1][3] += kk.y * va.w;
        acc[1][4] += kk.y * vb.x; acc[1][5] += kk.y * vb.y; acc[1][6] += kk.y * vb.z; acc[1][7] += kk.y * vb.w;
      }
    }
    __syncthreads();
  }

  // reduce the 4 pixel groups and write the partial
  for (int g = 0; g < 4; ++g) {
    if (pg4 == g) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int idx = (d0 + i) * 32 + eo * 8 + j;
          cred[idx] = (g == 0 ? 0.f : cred[idx]) + acc[i][j];
        }
    }
    __syncthreads();
  }
  float* out = partial + ((static_cast<int64_t>(b) * 4 + h) * nchunks + chunk) * kPartial;
  for (int i = t; i < 1024; i += 256) out[i] = cred[i];
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
  linattn_context_kernel<<<dim3(nchunks, 4, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(
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
