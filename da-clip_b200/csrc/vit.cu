// DA-CLIP image-encoder glue kernels (open_clip/transformer.py:507-555 of the reference): patch unfolding for
// the conv1-as-GEMM, class token + positional embedding + ln_pre, class-token pooling + ln_post + projection,
// and the degradation-type argmax (da-clip/src/evaluate_daclip.py:46-47,79-81).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/dac_b200.h"
#include "common.h"
#include "ptx.cuh"

namespace dac {

// out[(b*g*g + gy*g + gx), c*p*p + py*p + px] = image[b, c, gy*p+py, gx*p+px]
__global__ void __launch_bounds__(256) vit_patchify_kernel(const float* __restrict__ img,
                                                           __nv_bfloat16* __restrict__ out, int B, int S, int p,
                                                           int ld_out) {
  const int g = S / p;
  const int K = 3 * p * p;
  const int64_t total = static_cast<int64_t>(B) * g * g * K / 2;  // two px per thread
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t e = i * 2;
    const int k = static_cast<int>(e % K);
    const int64_t row = e / K;
    const int gx = static_cast<int>(row % g), gy = static_cast<int>((row / g) % g), b = static_cast<int>(row / (g * g));
    const int px = k % p, py = (k / p) % p, c = k / (p * p);
    const float* src = img + ((static_cast<int64_t>(b) * 3 + c) * S + gy * p + py) * S + gx * p + px;
    *reinterpret_cast<uint32_t*>(out + row * ld_out + k) = pack_bf16(src[0], src[1]);
  }
}

// p % 8 == 0 (ViT-B/32): eight pixels per thread - two 16-byte loads, one 16-byte store, one set of divisions per eight
// pixels (the two-pixel version above spends its time in integer divisions: 171 us for 154 MB at batch 256)
__global__ void __launch_bounds__(256) vit_patchify8_kernel(const float* __restrict__ img,
                                                            __nv_bfloat16* __restrict__ out, int B, int S, int p,
                                                            int ld_out) {
  const int g = S / p;
  const int pv = p >> 3;                               // 8-pixel vectors per patch row
  const int kv = 3 * p * pv;                           // ... per patch
  const int64_t total = static_cast<int64_t>(B) * g * g * kv;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int v = static_cast<int>(i % kv);
    const int64_t row = i / kv;
    const int gx = static_cast<int>(row % g), gy = static_cast<int>((row / g) % g), b = static_cast<int>(row / (g * g));
    const int x8 = v % pv, py = (v / pv) % p, c = v / (pv * p);
    const float4* src = reinterpret_cast<const float4*>(
        img + ((static_cast<int64_t>(b) * 3 + c) * S + gy * p + py) * S + gx * p + x8 * 8);
    const float4 a = __ldg(src), d = __ldg(src + 1);
    *reinterpret_cast<uint4*>(out + row * ld_out + (c * p + py) * p + x8 * 8) =
        make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(d.x, d.y), pack_bf16(d.z, d.w));
  }
}

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) t += red[i];
  return t;
}

// one CTA per token: (cls | patch embedding) + positional embedding -> LayerNorm -> bf16
__global__ void __launch_bounds__(256) vit_embed_kernel(const __nv_bfloat16* __restrict__ patch_emb,
                                                        const float* __restrict__ cls, const float* __restrict__ pos,
                                                        const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                        float* __restrict__ out, int L, int w, float eps) {
  extern __shared__ float row[];
  __shared__ float red[8];
  const int tok = blockIdx.x % L, b = blockIdx.x / L;
  float s = 0.f;
  for (int i = threadIdx.x; i < w; i += blockDim.x) {
    float v = (tok == 0) ? cls[i]
                         : __bfloat162float(patch_emb[(static_cast<int64_t>(b) * (L - 1) + tok - 1) * w + i]);
    v += pos[static_cast<int64_t>(tok) * w + i];
    row[i] = v;
    s += v;
  }
  const float mean = block_sum(s, red) / w;
  float ss = 0.f;
  for (int i = threadIdx.x; i < w; i += blockDim.x) {
    const float d = row[i] - mean;
    ss += d * d;
  }
  const float rstd = rsqrtf(block_sum(ss, red) / w + eps);
  for (int i = threadIdx.x; i < w; i += blockDim.x)
    out[(static_cast<int64_t>(b) * L + tok) * w + i] = (row[i] - mean) * rstd * ln_w[i] + ln_b[i];
}

// kPoolImgs images per CTA: ln_post(x[b,0,:]) for each, then @ proj with every projection row read once per CTA
constexpr int kPoolImgs = 8;
// (row_index: the token pooled per image - NULL = the class token 0; the text tower pools its end-of-text token)
__global__ void __launch_bounds__(256) vit_pool_kernel(const float* __restrict__ x, const int* __restrict__ row_index,
                                                       int B, int L, int w,
                                                       const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                       float eps, const float* __restrict__ proj, int e,
                                                       float* __restrict__ out) {
  extern __shared__ float rows[];   // [kPoolImgs][w]
  const int b0 = blockIdx.x * kPoolImgs;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < kPoolImgs && b0 + warp < B) {   // one warp normalises one class token
    const float* src = x + (static_cast<int64_t>(b0 + warp) * L + (row_index ? row_index[b0 + warp] : 0)) * w;
    float* row = rows + warp * w;
    float s = 0.f;
    for (int i = lane; i < w; i += 32) {
      row[i] = src[i];
      s += row[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / w;
    float ss = 0.f;
    for (int i = lane; i < w; i += 32) {
      const float d = row[i] - mean;
      ss += d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss / w + eps);
    for (int i = lane; i < w; i += 32) row[i] = (row[i] - mean) * rstd * ln_w[i] + ln_b[i];
  }
  __syncthreads();
  // blockIdx.y owns 64 projection columns; the 256 threads = 64 columns x 4 slices of the w contraction, reduced
  // through shared memory (fixed order: slice 0 + 1 + 2 + 3)
  __shared__ float part[4][kPoolImgs][64];
  const int jl = threadIdx.x & 63, ks = threadIdx.x >> 6;
  const int j = blockIdx.y * 64 + jl;
  const int i0 = ks * (w / 4), i1 = ks == 3 ? w : i0 + w / 4;
  float a[kPoolImgs];
#pragma unroll
  for (int m = 0; m < kPoolImgs; ++m) a[m] = 0.f;
  if (j < e) {
    for (int i = i0; i < i1; ++i) {
      const float pj = __ldg(proj + static_cast<int64_t>(i) * e + j);   // coalesced over j
#pragma unroll
      for (int m = 0; m < kPoolImgs; ++m) a[m] = fmaf(rows[m * w + i], pj, a[m]);
    }
  }
#pragma unroll
  for (int m = 0; m < kPoolImgs; ++m) part[ks][m][jl] = a[m];
  __syncthreads();
  if (ks == 0 && j < e) {
#pragma unroll
    for (int m = 0; m < kPoolImgs; ++m)
      if (b0 + m < B)
        out[static_cast<int64_t>(b0 + m) * e + j] = ((part[0][m][jl] + part[1][m][jl]) + part[2][m][jl]) + part[3][m][jl];
  }
}

// CLIP text-tower entry (open_clip/model.py:240-242,248): one CTA per token, x[b,t,:] = token_embedding[text[b,t]] +
// positional_embedding[t] in fp32; the CTA of token 0 also records eot[b] = argmax_t text[b,t] (first maximum, the
// torch.argmax rule; the end-of-text id is the largest of the vocabulary).  Ids outside [0, vocab) are clamped here;
// the host binding rejects them before the launch (the reference raises an IndexError).
__global__ void __launch_bounds__(128) text_embed_kernel(const int64_t* __restrict__ text,
                                                         const float* __restrict__ emb, const float* __restrict__ pos,
                                                         float* __restrict__ out, int* __restrict__ eot, int L, int w,
                                                         int vocab) {
  const int t = blockIdx.x % L, b = blockIdx.x / L;
  int64_t id = text[static_cast<int64_t>(b) * L + t];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const float4* e = reinterpret_cast<const float4*>(emb + id * w);
  const float4* p = reinterpret_cast<const float4*>(pos + static_cast<int64_t>(t) * w);
  float4* o = reinterpret_cast<float4*>(out + (static_cast<int64_t>(b) * L + t) * w);
  for (int i = threadIdx.x; i < w / 4; i += blockDim.x) {
    const float4 a = __ldg(e + i), c = __ldg(p + i);
    o[i] = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
  }
  if (t == 0 && threadIdx.x == 0) {
    int best = 0;
    int64_t bv = text[static_cast<int64_t>(b) * L];
    for (int i = 1; i < L; ++i) {
      const int64_t v = text[static_cast<int64_t>(b) * L + i];
      if (v > bv) {
        bv = v;
        best = i;
      }
    }
    eot[b] = best;
  }
}

// one warp per image: logits[j] = 100 * <d/|d|, t_j/|t_j|>; first index of the maximum (torch.argmax tie rule)
__global__ void __launch_bounds__(32) degradation_argmax_kernel(const float* __restrict__ degra,
                                                                const float* __restrict__ text, int e, int classes,
                                                                float* __restrict__ logits,
                                                                int64_t* __restrict__ argmax) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const float* d = degra + static_cast<int64_t>(b) * e;
  float dn = 0.f;
  for (int i = lane; i < e; i += 32) dn += d[i] * d[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dn += __shfl_xor_sync(0xffffffffu, dn, o);
  dn = sqrtf(dn);
  float best = -INFINITY;
  int besti = 0;
  for (int j = 0; j < classes; ++j) {
    const float* t = text + static_cast<int64_t>(j) * e;
    float tn = 0.f, dot = 0.f;
    for (int i = lane; i < e; i += 32) {
      tn += t[i] * t[i];
      dot += (d[i] / dn) * t[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      tn += __shfl_xor_sync(0xffffffffu, tn, o);
      dot += __shfl_xor_sync(0xffffffffu, dot, o);
    }
    const float lg = 100.0f * dot / sqrtf(tn);
    if (logits && lane == 0) logits[static_cast<int64_t>(b) * classes + j] = lg;
    if (lg > best) {
      best = lg;
      besti = j;
    }
  }
  if (lane == 0) argmax[b] = besti;
}

}  // namespace dac

using namespace dac;

extern "C" int dac_vit_patchify(const float* image, void* out, int32_t B, int32_t S, int32_t p, int32_t ld_out,
                                dac_stream_t stream) {
  if (!image || !out) return set_error(-1, "dac_vit_patchify: null argument");
  if (S % p || (p & 1)) return set_error(-2, "dac_vit_patchify: patch must be even and divide the image");
  if (ld_out < 3 * p * p || (ld_out & 1)) return set_error(-2, "dac_vit_patchify: ld_out must be even and >= 3*p*p");
  const int g = S / p;
  if (p % 8 == 0 && S % 8 == 0 && ld_out % 8 == 0 && !((reinterpret_cast<uintptr_t>(image) | reinterpret_cast<uintptr_t>(out)) & 15)) {
    int64_t blocks8 = ceil_div(static_cast<int64_t>(B) * g * g * 3 * p * (p / 8), 256);
    if (blocks8 > 148 * 32) blocks8 = 148 * 32;
    vit_patchify8_kernel<<<static_cast<int>(blocks8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        image, static_cast<__nv_bfloat16*>(out), B, S, p, ld_out);
    return check_launch("vit_patchify8_kernel");
  }
  const int64_t total = static_cast<int64_t>(B) * g * g * 3 * p * p / 2;
  int64_t blocks = ceil_div(total, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  vit_patchify_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      image, static_cast<__nv_bfloat16*>(out), B, S, p, ld_out);
  return check_launch("vit_patchify_kernel");
}

extern "C" int dac_vit_embed(const void* patch_emb, const float* cls, const float* pos, const float* ln_w,
                             const float* ln_b, void* out, int32_t B, int32_t L, int32_t w, float eps,
                             dac_stream_t stream) {
  if (!patch_emb || !cls || !pos || !ln_w || !ln_b || !out) return set_error(-1, "dac_vit_embed: null argument");
  vit_embed_kernel<<<B * L, 256, w * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(patch_emb), cls, pos, ln_w, ln_b, static_cast<float*>(out), L, w, eps);
  return check_launch("vit_embed_kernel");
}

extern "C" int dac_vit_pool(const void* x, int32_t B, int32_t L, int32_t w, const float* ln_w, const float* ln_b,
                            float eps, const float* proj, int32_t e, float* out, dac_stream_t stream) {
  if (!x || !ln_w || !ln_b || !proj || !out) return set_error(-1, "dac_vit_pool: null argument");
  vit_pool_kernel<<<dim3((B + kPoolImgs - 1) / kPoolImgs, (e + 63) / 64), 256, kPoolImgs * w * sizeof(float),
                    static_cast<cudaStream_t>(stream)>>>(static_cast<const float*>(x), nullptr, B, L, w, ln_w, ln_b, eps,
                                                         proj, e, out);
  return check_launch("vit_pool_kernel");
}

extern "C" int dac_text_embed(const int64_t* text, const float* token_embedding, const float* pos, float* out,
                              int32_t* eot, int32_t B, int32_t L, int32_t w, int32_t vocab, dac_stream_t stream) {
  if (!text || !token_embedding || !pos || !out || !eot) return set_error(-1, "dac_text_embed: null argument");
  if (B <= 0 || L <= 0 || w <= 0 || (w & 3) || vocab <= 0)
    return set_error(-2, "dac_text_embed: sizes must be positive, width a multiple of 4");
  text_embed_kernel<<<B * L, 128, 0, static_cast<cudaStream_t>(stream)>>>(text, token_embedding, pos, out, eot, L, w,
                                                                         vocab);
  return check_launch("text_embed_kernel");
}

extern "C" int dac_text_pool(const void* x, const int32_t* eot, int32_t B, int32_t L, int32_t w, const float* ln_w,
                             const float* ln_b, float eps, const float* proj, int32_t e, float* out,
                             dac_stream_t stream) {
  if (!x || !eot || !ln_w || !ln_b || !proj || !out) return set_error(-1, "dac_text_pool: null argument");
  vit_pool_kernel<<<dim3((B + kPoolImgs - 1) / kPoolImgs, (e + 63) / 64), 256, kPoolImgs * w * sizeof(float),
                    static_cast<cudaStream_t>(stream)>>>(static_cast<const float*>(x), eot, B, L, w, ln_w, ln_b, eps,
                                                         proj, e, out);
  return check_launch("vit_pool_kernel(text)");
}

extern "C" int dac_degradation_argmax(const float* degra, const float* text, int32_t B, int32_t e, int32_t classes,
                                      float* logits, int64_t* argmax, dac_stream_t stream) {
  if (!degra || !text || !argmax) return set_error(-1, "dac_degradation_argmax: null argument");
  degradation_argmax_kernel<<<B, 32, 0, static_cast<cudaStream_t>(stream)>>>(degra, text, e, classes, logits, argmax);
  return check_launch("degradation_argmax_kernel");
}
