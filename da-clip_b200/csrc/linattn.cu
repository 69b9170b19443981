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
#include "linattn_kv_common.h"
#include "ptx.cuh"
#include "tile_common.cuh"

namespace dac {

constexpr int kLaP = 64;                 // pixels per sub-tile
constexpr int kLaPitch = kLaP + 8;       // bf16 pitch of the transposed [channel][pixel] tiles
constexpr int kPartial = 32 * 32 + 64;

// One CTA (4 warps) per (pixel slab, head, image).  k and v arrive PLANAR ([B][256][hw], pixel-contiguous channel
// rows - written that way by the QKV epilogue), so a (channel, 16-pixel quarter) thread loads its 32 bytes of k and
// of v straight from global memory: running column max / rescale / exp / row sums on k, then both go to shared
// memory as [channel][pixel] bf16 tiles - the operand layout of C[d][e] += sum_p exp(k)[d][p] v[e][p] on the
// warp-level tensor cores (mma.sync m16n8k16 bf16, fp32 accumulate): warp w owns d rows 16*(w/2).., e cols 16*(w%2)..
__global__ void __launch_bounds__(128) linattn_context_kernel(const __nv_bfloat16* __restrict__ kv, int hw,
                                                              int slab, float* __restrict__ partial) {
  __shared__ __align__(16) __nv_bfloat16 Kt[32 * kLaPitch];
  __shared__ __align__(16) __nv_bfloat16 Vt[32 * kLaPitch];
  __shared__ float m_run[32], s_run[32], scale[32];

  const int chunk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nchunks = gridDim.x;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int p_begin = chunk * slab, p_end = min(hw, p_begin + slab);
  const int d = t >> 2, qtr = t & 3;      // load / softmax mapping: channel row d, 16-pixel quarter
  const __nv_bfloat16* krow = kv + (static_cast<int64_t>(b) * 256 + h * 32 + d) * hw;
  const __nv_bfloat16* vrow = kv + (static_cast<int64_t>(b) * 256 + 128 + h * 32 + d) * hw;
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
  __syncthreads();

  for (int p0 = p_begin; p0 < p_end; p0 += kLaP) {
    const int px = p0 + qtr * 16;
    float kf[16];
    uint4 uv[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
    if (px + 16 <= p_end) {
      const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(krow + px));
      const uint4 u1 = __ldg(reinterpret_cast<const uint4*>(krow + px) + 1);
      uv[0] = __ldg(reinterpret_cast<const uint4*>(vrow + px));
      uv[1] = __ldg(reinterpret_cast<const uint4*>(vrow + px) + 1);
      const uint32_t w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 f = unpack_bf16(w[j]);
        kf[2 * j] = f.x;
        kf[2 * j + 1] = f.y;
      }
    } else {  // ragged tail: element-wise, masked pixels contribute exp(-inf) = 0 and v = 0
      __nv_bfloat16* vb = reinterpret_cast<__nv_bfloat16*>(uv);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const bool live = px + j < p_end;
        kf[j] = live ? __bfloat162float(krow[px + j]) : -INFINITY;
        vb[j] = live ? vrow[px + j] : __float2bfloat16(0.f);
      }
    }
    float m = kf[0];
#pragma unroll
    for (int j = 1; j < 16; ++j) m = fmaxf(m, kf[j]);
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    __syncthreads();                       // previous sub-tile's MMAs are done: Kt / Vt / scale may be rewritten
    const float mo = m_run[d];
    const float mn = fmaxf(mo, m);         // finite: the first sub-tile of a slab holds at least one live pixel
    float s = 0.f;
    uint32_t packed[8];
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      packed[j >> 1] = pack_bf16(__expf(kf[j] - mn), __expf(kf[j + 1] - mn));
      const float2 r = unpack_bf16(packed[j >> 1]);        // sum exactly what the tensor cores will multiply
      s += r.x + r.y;
    }
    __nv_bfloat16* kd = &Kt[d * kLaPitch + qtr * 16];
    *reinterpret_cast<uint4*>(kd) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    *reinterpret_cast<uint4*>(kd + 8) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
    __nv_bfloat16* vd = &Vt[d * kLaPitch + qtr * 16];
    *reinterpret_cast<uint4*>(vd) = uv[0];
    *reinterpret_cast<uint4*>(vd + 8) = uv[1];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    __syncwarp();
    if (qtr == 0) {
      const float sc = (mo == -INFINITY) ? 0.f : __expf(mo - mn);
      scale[d] = sc;
      m_run[d] = mn;
      s_run[d] = s_run[d] * sc + s;
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

// One CTA per (head, image, group of kFoldRows d rows): the merge is a sum over `nchunks` partial records in a FIXED
// order (two interleaved chains + one add; warp-shuffle trees for the denominators), so the result is bit-reproducible
// whichever CTA of the producing kernel finished first.  Splitting the d rows over CTAs keeps the kernel short when one
// image spans every CTA of the producer (batch 1: 148 or 296 records per head).
constexpr int kFoldRows = 4;
__global__ void __launch_bounds__(256) linattn_fold_kernel(const float* __restrict__ partial, int hw, int nchunks,
                                                           const float* __restrict__ w_out, int C, int c_pad,
                                                           __nv_bfloat16* __restrict__ weff) {
  extern __shared__ float fold_sm[];
  float* wsm = fold_sm;                          // [C][33]: W_out[c][h*32 + e]
  float* wgt = wsm + C * 33;                     // [nchunks][kFoldRows] merge weights exp(m_c - M)
  float* red = wgt + nchunks * kFoldRows;        // [2][kFoldRows * 32] the two chains
  float* ctx = red + 2 * kFoldRows * 32;         // [kFoldRows][33]
  __shared__ float inv_s[kFoldRows];
  const int h = blockIdx.x, b = blockIdx.y, d0 = blockIdx.z * kFoldRows, t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  const float* pbase = partial + (static_cast<int64_t>(b) * 4 + h) * nchunks * kPartial;
  griddep_launch();
  for (int i = t; i < C * 32; i += 256) {       // a constant: staged while the context kernel drains
    const int c = i >> 5, e = i & 31;
    wsm[c * 33 + e] = __ldg(w_out + static_cast<int64_t>(c) * 128 + h * 32 + e);
  }
  griddep_wait();
  if (warp < kFoldRows) {                        // warp w: running-max merge weights and the denominator of row d0 + w
    const int d = d0 + warp;
    float M = -INFINITY;
    for (int c = lane; c < nchunks; c += 32) M = fmaxf(M, pbase[c * kPartial + 1024 + d]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
    float S = 0.f;
    for (int c = lane; c < nchunks; c += 32) {
      const float mc = pbase[c * kPartial + 1024 + d];
      const float w = (mc == -INFINITY) ? 0.f : __expf(mc - M);
      wgt[c * kFoldRows + warp] = w;
      S += pbase[c * kPartial + 1056 + d] * w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) S += __shfl_xor_sync(0xffffffffu, S, o);
    if (lane == 0) inv_s[warp] = 1.0f / (S * static_cast<float>(hw));   // softmax denominator and v / (h*w) (module_util.py:177)
  }
  __syncthreads();
  {
    const int elem = t & (kFoldRows * 32 - 1), part = t >> 7;            // 128 context elements x 2 chains
    const int dl = elem >> 5;
    const float* src = pbase + d0 * 32 + elem;
    float a = 0.f;
#pragma unroll 4
    for (int c = part; c < nchunks; c += 2) a += src[static_cast<int64_t>(c) * kPartial] * wgt[c * kFoldRows + dl];
    red[part * (kFoldRows * 32) + elem] = a;
  }
  __syncthreads();
  if (t < kFoldRows * 32) ctx[(t >> 5) * 33 + (t & 31)] = (red[t] + red[kFoldRows * 32 + t]) * inv_s[t >> 5];
  __syncthreads();
  const int dl = t & (kFoldRows - 1);
  for (int c = t >> 2; c < C; c += 64) {
    float a = 0.f;
#pragma unroll
    for (int e = 0; e < 32; ++e) a = fmaf(wsm[c * 33 + e], ctx[dl * 33 + e], a);
    weff[(static_cast<int64_t>(b) * c_pad + c) * 128 + h * 32 + d0 + dl] = __float2bfloat16(a);
  }
}

// The fold for dac_linattn_kv's in-kernel-PreNorm mode (linattn_kv2.cu), whose partial records hold G = P^T xn
// ([32 d][64 c]) and S instead of the context: context = G W_v^T / (S hw) and W_eff = W_out context^T collapse into
//   weff[b][c'][h*32+d] = sum_c G[b,h,d,c] M_h[c'][c] / (S[b,h,d] hw),   M_h = W_out[:, h] W_v[h]   (constant, fp32)
// One CTA per (head, image, group of kFoldRows d rows); slots are added in a FIXED order (two interleaved chains + one
// add, shuffle trees for S): bit-reproducible.  No running-max weights: the k|v kernels use a data-independent shift.
__global__ void __launch_bounds__(256) linattn_fold_g_kernel(const float* __restrict__ partial, int hw, int slots,
                                                             int tiles, int grid, const float* __restrict__ m_fold, int C,
                                                             int c_pad, __nv_bfloat16* __restrict__ weff) {
  extern __shared__ float fold_sm[];
  float* msm = fold_sm;                          // [C][65]: M_h[c'][c]
  float* g = msm + C * 65;                       // [kFoldRows][64]
  __shared__ float inv_s[kFoldRows];
  const int h = blockIdx.x, b = blockIdx.y, d0 = blockIdx.z * kFoldRows, t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  const float* pbase = partial + (static_cast<int64_t>(b) * 4 + h) * slots * kKvGRec;
  // the records that exist: one per CTA of the producing kernel whose tile range touches image b (the rest of the
  // `slots` records is never written - and never read, so ctx_acc needs no clearing between launches)
  const int tpi = hw / 128;
  const int nslots = tile_owner((b + 1) * tpi - 1, tiles, grid) - tile_owner(b * tpi, tiles, grid) + 1;
  griddep_launch();
  for (int i = t; i < C * 64; i += 256)         // a constant: staged while the producing kernel drains
    msm[(i >> 6) * 65 + (i & 63)] = __ldg(m_fold + static_cast<int64_t>(h) * C * 64 + i);
  griddep_wait();
  if (warp < kFoldRows) {
    float S = 0.f;
    for (int c = lane; c < nslots; c += 32) S += pbase[static_cast<int64_t>(c) * kKvGRec + 2048 + d0 + warp];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) S += __shfl_xor_sync(0xffffffffu, S, o);
    if (lane == 0) inv_s[warp] = 1.0f / (S * static_cast<float>(hw));   // softmax denominator and v / (h*w) (module_util.py:177)
  }
  float a0 = 0.f, a1 = 0.f;
  {
    const float* src = pbase + d0 * 64 + t;      // 4 rows x 64 channels = 256 consecutive floats of a record
    int c = 0;
#pragma unroll 4
    for (; c + 1 < nslots; c += 2) {      // (eight independent loads in flight: the merge is one L2 round trip per 8 slots)
      a0 += src[static_cast<int64_t>(c) * kKvGRec];
      a1 += src[static_cast<int64_t>(c + 1) * kKvGRec];
    }
    if (c < nslots) a0 += src[static_cast<int64_t>(c) * kKvGRec];
  }
  __syncthreads();
  g[t] = (a0 + a1) * inv_s[t >> 6];
  __syncthreads();
  const int dl = t & (kFoldRows - 1);
  for (int c = t >> 2; c < C; c += 64) {
    float a = 0.f;
#pragma unroll 16
    for (int e = 0; e < 64; ++e) a = fmaf(msm[c * 65 + e], g[dl * 64 + e], a);
    weff[(static_cast<int64_t>(b) * c_pad + c) * 128 + h * 32 + d0 + dl] = __float2bfloat16(a);
  }
}

}  // namespace dac

using namespace dac;

extern "C" int dac_linattn_context(const void* kv, int32_t B, int32_t hw, int32_t nchunks, float* partial,
                                   dac_stream_t stream) {
  if (!kv || !partial) return set_error(-1, "dac_linattn_context: null argument");
  if (hw & 7) return set_error(-2, "dac_linattn_context: hw must be a multiple of 8");
  if (nchunks < 1 || nchunks > 128) return set_error(-2, "dac_linattn_context: nchunks must be in [1,128]");
  const int slab = static_cast<int>(ceil_div(ceil_div(hw, nchunks), kLaP) * kLaP);
  linattn_context_kernel<<<dim3(nchunks, 4, B), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(kv), hw, slab, partial);
  return check_launch("linattn_context_kernel");
}

extern "C" int dac_linattn_fold(const float* partial, int32_t B, int32_t hw, int32_t nchunks, const float* w_out,
                                int32_t C, int32_t c_pad, void* weff, dac_stream_t stream) {
  if (!partial || !w_out || !weff) return set_error(-1, "dac_linattn_fold: null argument");
  if (nchunks < 1 || nchunks > 2048) return set_error(-2, "dac_linattn_fold: nchunks must be in [1,2048]");
  if (C <= 0 || C > 1024) return set_error(-2, "dac_linattn_fold: C must be in [1,1024]");
  const size_t smem = sizeof(float) * (static_cast<size_t>(C) * 33 + static_cast<size_t>(nchunks) * kFoldRows +
                                       2 * kFoldRows * 32 + kFoldRows * 33);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(linattn_fold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(-12, "dac_linattn_fold: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  launch_k(linattn_fold_kernel, dim3(4, B, 32 / kFoldRows), dim3(256), smem, static_cast<cudaStream_t>(stream), partial, hw, nchunks, w_out,
           C, c_pad, static_cast<__nv_bfloat16*>(weff));
  return check_launch("linattn_fold_kernel");
}

extern "C" int dac_linattn_fold_g(const float* partial, int32_t B, int32_t hw, int32_t nslots, const float* m_fold,
                                  int32_t C, int32_t c_pad, void* weff, dac_stream_t stream) {
  if (!partial || !m_fold || !weff) return set_error(-1, "dac_linattn_fold_g: null argument");
  if (nslots < 1 || nslots > 2048) return set_error(-2, "dac_linattn_fold_g: nslots must be in [1,2048]");
  if (C <= 0 || C > 128 || c_pad < C) return set_error(-2, "dac_linattn_fold_g: C must be in [1,128], c_pad >= C");
  if (hw <= 0 || hw % 128) return set_error(-2, "dac_linattn_fold_g: hw must be a multiple of 128");
  const size_t smem = sizeof(float) * (static_cast<size_t>(C) * 65 + kFoldRows * 64);
  // the producer's geometry (dac_linattn_kv_create): B * hw / 128 tiles in contiguous ranges over min(tiles, SMs) CTAs
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = B * (hw / 128), grid = tiles < sms ? tiles : sms;
  if (nslots < max_image_span(B, hw / 128, grid)) return set_error(-2, "dac_linattn_fold_g: nslots below dac_linattn_ctx_slots");
  launch_k(linattn_fold_g_kernel, dim3(4, B, 32 / kFoldRows), dim3(256), smem, static_cast<cudaStream_t>(stream), partial, hw,
           nslots, tiles, grid, m_fold, C, c_pad, static_cast<__nv_bfloat16*>(weff));
  return check_launch("linattn_fold_g_kernel");
}
