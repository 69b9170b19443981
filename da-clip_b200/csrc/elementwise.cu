// HBM-bound kernels of the path: SDE / posterior / ODE state update, stem input packing, row LayerNorm,
// GroupNorm, and the tiny fp32 conditioning MLPs (time / prompt embedding -> FiLM table).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <curand_kernel.h>

#include "../../include/dac_b200.h"
#include "common.h"
#include "ptx.cuh"

namespace dac {

// ------------------------------------------------------------------------------------------------ SDE step
// Arithmetic is spelled with non-contracting intrinsics in the reference's exact operation order
// (sde_utils.py:44-45,177-187,205-231,245-247) so the result is bit-identical to the fp32 PyTorch expression.
struct SdeCoef {
  float c[8];
};

template <int MODE>
__device__ __forceinline__ float sde_update(float x, float mu, float n, float e, const SdeCoef& k) {
  if (MODE == 0 || MODE == 2) {
    // c: 0 theta, 1 sigma^2 (or 0.5*sigma^2 for the ODE), 2 sigma_bar, 3 dt, 4 sigma, 5 sqrt(dt)
    const float score = __fdiv_rn(-n, k.c[2]);
    const float drift = __fmul_rn(__fsub_rn(__fmul_rn(k.c[0], __fsub_rn(mu, x)), __fmul_rn(k.c[1], score)), k.c[3]);
    float r = __fsub_rn(x, drift);
    if (MODE == 0) r = __fsub_rn(r, __fmul_rn(k.c[4], __fmul_rn(e, k.c[5])));
    return r;
  } else {
    // c: 0 term1, 1 term2, 2 std, 3 exp(Theta_t dt), 4 sigma_bar
    const float xm = __fsub_rn(x, mu);
    const float x0 = __fadd_rn(__fmul_rn(__fsub_rn(xm, __fmul_rn(k.c[4], n)), k.c[3]), mu);
    const float mean = __fadd_rn(__fadd_rn(__fmul_rn(k.c[0], xm), __fmul_rn(k.c[1], __fsub_rn(x0, mu))), mu);
    return __fadd_rn(mean, __fmul_rn(k.c[2], e));
  }
}

template <int MODE>
// (x and out carry no __restrict__: the samplers update the state in place, out == x; every element is read before it is
// written by the same thread)
__global__ void __launch_bounds__(256) sde_step_kernel(const float* x, const float* __restrict__ mu,
                                                       const float* __restrict__ net, const float* __restrict__ eps,
                                                       float* out, int64_t n, SdeCoef k) {
  const int64_t n4 = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = reinterpret_cast<const float4*>(x)[i];
    const float4 m = reinterpret_cast<const float4*>(mu)[i];
    const float4 nn = reinterpret_cast<const float4*>(net)[i];
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODE != 2) e = reinterpret_cast<const float4*>(eps)[i];
    float4 r;
    r.x = sde_update<MODE>(a.x, m.x, nn.x, e.x, k);
    r.y = sde_update<MODE>(a.y, m.y, nn.y, e.y, k);
    r.z = sde_update<MODE>(a.z, m.z, nn.z, e.z, k);
    r.w = sde_update<MODE>(a.w, m.w, nn.w, e.w, k);
    reinterpret_cast<float4*>(out)[i] = r;
  }
  // tail (n not a multiple of 4)
  for (int64_t i = (n4 << 2) + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = sde_update<MODE>(x[i], mu[i], net[i], MODE != 2 ? eps[i] : 0.f, k);
}

// ---- the sampling loop driven from the device: no per-step host work, no per-step ATen launch ----
// loop_tick_kernel (first node of a step, one thread): step s = state[0]; publishes the network time of this step in
// t_dev (what time_embed reads), its 8 update coefficients in coef_dev, s itself in state[1], and advances state[0].
__global__ void loop_tick_kernel(long long* __restrict__ state, const float* __restrict__ t_table,
                                 const float* __restrict__ coef_table, float* __restrict__ t_dev,
                                 float* __restrict__ coef_dev) {
  if (threadIdx.x == 0) {
    const long long s = state[0];
    t_dev[0] = t_table[s];
    state[1] = s;
    state[0] = s + 1;
  }
  __syncthreads();
  if (threadIdx.x < 8) coef_dev[threadIdx.x] = coef_table[state[1] * 8 + threadIdx.x];
}

// sde_step_kernel with the coefficients and the step index read from device memory (written by loop_tick_kernel of the
// same step) and the Gaussian noise either read from a pre-generated [T][n] tensor (eps_base + step * n: the parity
// tests inject the reference's draws) or generated here: Philox4x32-10 keyed by (seed, element quad, step), four normals
// per call - the sampler then launches nothing but its own kernels (sde_utils.py:227-231 draws torch.randn_like per step).
template <int MODE>
__global__ void __launch_bounds__(256) sde_step_dev_kernel(const float* x, const float* __restrict__ mu,
                                                           const float* __restrict__ net, float* out,
                                                           int64_t n, const float* __restrict__ coef_dev,
                                                           const long long* __restrict__ state) {
  griddep_wait();      // launched with programmatic serialisation behind final_conv: its output must be complete
  SdeCoef k;
#pragma unroll
  for (int i = 0; i < 8; ++i) k.c[i] = coef_dev[i];
  const long long step = state[1];
  const float* eps_base = reinterpret_cast<const float*>(state[2]);   // the host sets it before the first step
  const unsigned long long seed = static_cast<unsigned long long>(state[3]);
  const float* eps = eps_base ? eps_base + step * n : nullptr;
  const int64_t n4 = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = reinterpret_cast<const float4*>(x)[i];
    const float4 m = reinterpret_cast<const float4*>(mu)[i];
    const float4 nn = reinterpret_cast<const float4*>(net)[i];
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODE != 2) {
      if (eps) {
        e = reinterpret_cast<const float4*>(eps)[i];
      } else {
        curandStatePhilox4_32_10_t st;
        curand_init(seed, static_cast<unsigned long long>(i), static_cast<unsigned long long>(step), &st);
        e = curand_normal4(&st);
      }
    }
    float4 r;
    r.x = sde_update<MODE>(a.x, m.x, nn.x, e.x, k);
    r.y = sde_update<MODE>(a.y, m.y, nn.y, e.y, k);
    r.z = sde_update<MODE>(a.z, m.z, nn.z, e.z, k);
    r.w = sde_update<MODE>(a.w, m.w, nn.w, e.w, k);
    reinterpret_cast<float4*>(out)[i] = r;
  }
}

__global__ void __launch_bounds__(256) noise_state_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                                                          float* __restrict__ out, int64_t n, float s) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = __fadd_rn(x[i], __fmul_rn(eps[i], s));
}

static int elementwise_grid(int64_t work_items) {
  int64_t blocks = ceil_div(work_items, 256);
  const int64_t cap = 148 * 16;
  return static_cast<int>(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

// ------------------------------------------------------------------------------------------------ stem input
// PAIR = false: one packed 64-channel row per pixel (kx = 0..6, 8 channels each; kx = 7 zero) -> out [B][Hp][Wp][64].
// PAIR = true: one packed row per PAIR of horizontally adjacent pixels, the 8-wide window kx' = 0..7 around them (source
// x = 2 j + kx' - 3): pixel 2j uses kx' 0..6, pixel 2j+1 uses kx' 1..7 -> out [B][Hp][Wp/2][64].  With the weights
// packed as 128 rows (ops.pack_stem_pair: the second pixel's copies shifted by one kx) init_conv becomes a 7-tap vertical
// conv with N = 128 over HALF as many GEMM rows: half the packed tensor, half the K steps per pixel, and MMAs that are
// not capped by the N = 64 operand fetch.
template <bool PAIR>
__global__ void __launch_bounds__(256) stem_input_kernel(const float* __restrict__ xt, const float* __restrict__ cond,
                                                         __nv_bfloat16* __restrict__ out, int B, int H, int W, int Hp,
                                                         int Wp) {
  // One CTA per (row, image).  Phase 1: the row's six channels (xt - cond | cond), reflect-padded to Wp and with the
  // three zero columns of the 7-tap window on either side, go to shared memory with coalesced loads.  Phase 2: one
  // thread per (pixel or pair, kx) packs its 8 channels (6 + 2 zeros) into one 16 B store - the row is written fully
  // coalesced and every input value is read from global memory once, not seven times.
  extern __shared__ float srow[];                      // [6][Wp + 6]; column j holds source x = j - 3
  const int y = blockIdx.x, b = blockIdx.y;
  const int pitch = Wp + 6;
  const int ys = y < H ? y : 2 * H - 2 - y;            // reflect (arch.py:111-116)
  const int64_t plane = static_cast<int64_t>(H) * W;
  const int64_t base = (static_cast<int64_t>(b) * 3) * plane + static_cast<int64_t>(ys) * W;
  griddep_launch();
  for (int j = threadIdx.x; j < pitch; j += blockDim.x) {
    const int xs = j - 3;
    float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (xs >= 0 && xs < Wp) {
      const int xr = xs < W ? xs : 2 * W - 2 - xs;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float cv = __ldg(cond + base + c * plane + xr);
        v[c] = __fsub_rn(__ldg(xt + base + c * plane + xr), cv);
        v[3 + c] = cv;
      }
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) srow[c * pitch + j] = v[c];
  }
  __syncthreads();
  const int units = PAIR ? Wp / 2 : Wp;                // packed rows of this image row
  uint4* orow = reinterpret_cast<uint4*>(out) + (static_cast<int64_t>(b) * Hp + y) * units * 8;
  for (int i = threadIdx.x; i < units * 8; i += blockDim.x) {
    const int kx = i & 7;
    const float* s = srow + (PAIR ? 2 : 1) * (i >> 3) + kx;   // source x = first pixel + kx - 3  ->  column first pixel + kx
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (PAIR || kx < 7) {
      u.x = pack_bf16(s[0], s[pitch]);
      u.y = pack_bf16(s[2 * pitch], s[3 * pitch]);
      u.z = pack_bf16(s[4 * pitch], s[5 * pitch]);
    }
    orow[i] = u;
  }
}

// ------------------------------------------------------------------------------------------------ row LayerNorm
// LPR lanes cooperate on one row (8 lanes for 64 channels ... 32 lanes for >= 256), so a warp covers 32/LPR rows and
// every lane moves 16 B vectors; the row is read ONCE and kept in registers (c <= LPR * 8 * VPL).
template <int LPR, int VPL>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const __nv_bfloat16* __restrict__ in, int ld_in,
                                                             __nv_bfloat16* __restrict__ out, int ld_out, int64_t rows,
                                                             int c, const float* __restrict__ w,
                                                             const float* __restrict__ b, float eps,
                                                             const __nv_bfloat16* __restrict__ res, int ld_res) {
  griddep_wait();       // programmatic dependent launch (common.h): nothing to do before the producer is done
  griddep_launch();
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const int rows_per_warp = 32 / LPR;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int nvec = c >> 3;
  const float inv_c = 1.0f / c;
  // gain / bias of this lane's vectors: loaded once per warp (16-byte parameter loads; per-row reloads were four loads per
  // data load), kept in registers across the warp's rows
  float wreg[VPL][8], breg[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = sub + i * LPR;
    float4 w0 = make_float4(1.f, 1.f, 1.f, 1.f), w1 = w0, b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
    if (w && vi < nvec) {
      w0 = __ldg(reinterpret_cast<const float4*>(w) + vi * 2);
      w1 = __ldg(reinterpret_cast<const float4*>(w) + vi * 2 + 1);
    }
    if (b && vi < nvec) {
      b0 = __ldg(reinterpret_cast<const float4*>(b) + vi * 2);
      b1 = __ldg(reinterpret_cast<const float4*>(b) + vi * 2 + 1);
    }
    wreg[i][0] = w0.x; wreg[i][1] = w0.y; wreg[i][2] = w0.z; wreg[i][3] = w0.w;
    wreg[i][4] = w1.x; wreg[i][5] = w1.y; wreg[i][6] = w1.z; wreg[i][7] = w1.w;
    breg[i][0] = b0.x; breg[i][1] = b0.y; breg[i][2] = b0.z; breg[i][3] = b0.w;
    breg[i][4] = b1.x; breg[i][5] = b1.y; breg[i][6] = b1.z; breg[i][7] = b1.w;
  }
  for (int64_t row0 = warp_global * rows_per_warp; row0 < rows; row0 += nwarps * rows_per_warp) {
    const int64_t row = row0 + lane / LPR;
    const bool live = row < rows;
    float v[VPL][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = sub + i * LPR;
      uint4 u = make_uint4(0, 0, 0, 0);
      if (live && vi < nvec) u = __ldg(reinterpret_cast<const uint4*>(in + row * ld_in) + vi);
      float2 t;
      t = unpack_bf16(u.x); v[i][0] = t.x; v[i][1] = t.y;
      t = unpack_bf16(u.y); v[i][2] = t.x; v[i][3] = t.y;
      t = unpack_bf16(u.z); v[i][4] = t.x; v[i][5] = t.y;
      t = unpack_bf16(u.w); v[i][6] = t.x; v[i][7] = t.y;
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[i][j];
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * inv_c;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (sub + i * LPR < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[i][j] - mean;
          ss = fmaf(d, d, ss);
        }
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss * inv_c + eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = sub + i * LPR;
      if (live && vi < nvec) {
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = fmaf((v[i][j] - mean) * rstd, wreg[i][j], breg[i][j]);
        if (res) {    // + residual row (the Residual wrapper around a wide LinearAttention, module_util.py:27-33)
          const uint4 r = __ldg(reinterpret_cast<const uint4*>(res + row * ld_res) + vi);
          float2 t;
          t = unpack_bf16(r.x); y[0] += t.x; y[1] += t.y;
          t = unpack_bf16(r.y); y[2] += t.x; y[3] += t.y;
          t = unpack_bf16(r.z); y[4] += t.x; y[5] += t.y;
          t = unpack_bf16(r.w); y[6] += t.x; y[7] += t.y;
        }
        uint4 o;
        o.x = pack_bf16(y[0], y[1]); o.y = pack_bf16(y[2], y[3]);
        o.z = pack_bf16(y[4], y[5]); o.w = pack_bf16(y[6], y[7]);
        reinterpret_cast<uint4*>(out + row * ld_out)[vi] = o;
      }
    }
  }
}

// fp32 rows in (ViT residual stream), bf16 rows out.  One warp per row; the row (NV float4 per lane, c <= 128 NV) is read
// ONCE into registers - the first version walked it three times (sum, centred squares, output), each pass a dependent trip
// to L2.  Persistent warps: weight and bias (the same 2 x NV float4 for every row) are loaded once per warp and stay in
// registers, so a row costs NV loads + NV stores instead of 3 NV loads + NV stores (16 -> 12 us at 12800 x 768).
template <int NV>
__global__ void __launch_bounds__(256) layernorm_rows_f32_kernel(const float* __restrict__ in, int ld_in,
                                                                 __nv_bfloat16* __restrict__ out, int ld_out,
                                                                 int64_t rows, int c, const float* __restrict__ w,
                                                                 const float* __restrict__ b, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int nvec = c >> 2;
  float4 wv[NV], bv[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = lane + 32 * k;
    wv[k] = (w && i < nvec) ? __ldg(reinterpret_cast<const float4*>(w) + i) : make_float4(1.f, 1.f, 1.f, 1.f);
    bv[k] = (b && i < nvec) ? __ldg(reinterpret_cast<const float4*>(b) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float inv_c = 1.0f / c;
  for (int64_t row = warp_global; row < rows; row += nwarps) {
    const float4* src = reinterpret_cast<const float4*>(in + row * ld_in);
    float4 u[NV];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = lane + 32 * k;
      u[k] = i < nvec ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      sum += (u[k].x + u[k].y) + (u[k].z + u[k].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * inv_c;
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      if (lane + 32 * k < nvec) {
        const float e0 = u[k].x - mean, e1 = u[k].y - mean, e2 = u[k].z - mean, e3 = u[k].w - mean;
        ss += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss * inv_c + eps);
    uint2* dst = reinterpret_cast<uint2*>(out + row * ld_out);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = lane + 32 * k;
      if (i < nvec) {
        const float v0 = (u[k].x - mean) * rstd * wv[k].x + bv[k].x;
        const float v1 = (u[k].y - mean) * rstd * wv[k].y + bv[k].y;
        const float v2 = (u[k].z - mean) * rstd * wv[k].z + bv[k].z;
        const float v3 = (u[k].w - mean) * rstd * wv[k].w + bv[k].w;
        dst[i] = make_uint2(pack_bf16(v0, v1), pack_bf16(v2, v3));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ GroupNorm
// Pass 1: CTA (slab, image) -> stats[b][slab][g] = {sum, sumsq} of its pixel slab.  Each thread owns one 8-channel vector
// column (8 | channels-per-group) and walks its pixels; the per-thread partials meet in shared memory and ONE thread per
// group adds them in a fixed order - no atomics anywhere, so the statistics are bit-reproducible.
constexpr int kGnSlabs = 16;
__global__ void __launch_bounds__(256) groupnorm_stats_kernel(const __nv_bfloat16* __restrict__ in, int hw, int c,
                                                              int groups, int slab, float* __restrict__ stats) {
  __shared__ float2 part[256];
  griddep_wait();
  griddep_launch();
  const int b = blockIdx.y;
  const int nvec = c >> 3;
  const int cpg = c / groups;
  const int p0 = blockIdx.x * slab, p1 = min(hw, p0 + slab);
  const int vec = threadIdx.x % nvec;
  const int prow = threadIdx.x / nvec;
  const int prows = blockDim.x / nvec;
  float s = 0.f, ss = 0.f;
  if (prow < prows) {
    for (int p = p0 + prow; p < p1; p += prows) {
      const uint4 u = *reinterpret_cast<const uint4*>(in + (static_cast<int64_t>(b) * hw + p) * c + vec * 8);
      const float2 a = unpack_bf16(u.x), bb = unpack_bf16(u.y), cc = unpack_bf16(u.z), d = unpack_bf16(u.w);
      s += (a.x + a.y) + (bb.x + bb.y) + (cc.x + cc.y) + (d.x + d.y);
      ss += a.x * a.x + a.y * a.y + bb.x * bb.x + bb.y * bb.y + cc.x * cc.x + cc.y * cc.y + d.x * d.x + d.y * d.y;
    }
  }
  part[threadIdx.x] = make_float2(s, ss);
  __syncthreads();
  if (threadIdx.x < groups) {
    const int g = threadIdx.x, vpg = cpg >> 3;
    float gs = 0.f, gss = 0.f;
    for (int r = 0; r < prows; ++r)
      for (int v = 0; v < vpg; ++v) {
        const float2 t = part[r * nvec + g * vpg + v];
        gs += t.x;
        gss += t.y;
      }
    reinterpret_cast<float2*>(stats)[(static_cast<int64_t>(b) * gridDim.x + blockIdx.x) * groups + g] = make_float2(gs, gss);
  }
}

// Pass 2: CTA (chunk, image): the image's group statistics are summed over the slabs in order, once per CTA.
__global__ void __launch_bounds__(256) groupnorm_apply_kernel(const __nv_bfloat16* __restrict__ in,
                                                              __nv_bfloat16* __restrict__ out, int hw, int c,
                                                              int groups, int slabs, const float* __restrict__ w,
                                                              const float* __restrict__ bias, float eps,
                                                              const float* __restrict__ stats) {
  __shared__ float2 mr[256];                     // {mean, rstd} per group
  griddep_wait();
  griddep_launch();
  const int b = blockIdx.y;
  const int nvec = c >> 3;
  const int cpg = c / groups;
  if (threadIdx.x < groups) {
    float s = 0.f, ss = 0.f;
    for (int k = 0; k < slabs; ++k) {
      const float2 t = __ldg(reinterpret_cast<const float2*>(stats) + (static_cast<int64_t>(b) * slabs + k) * groups + threadIdx.x);
      s += t.x;
      ss += t.y;
    }
    const float inv_n = 1.0f / (static_cast<float>(hw) * cpg);
    const float mean = s * inv_n;
    const float var = fmaxf(ss * inv_n - mean * mean, 0.f);
    mr[threadIdx.x] = make_float2(mean, rsqrtf(var + eps));
  }
  __syncthreads();
  const int64_t per_image = static_cast<int64_t>(hw) * nvec;
  const uint4* ib = reinterpret_cast<const uint4*>(in) + b * per_image;
  uint4* ob = reinterpret_cast<uint4*>(out) + b * per_image;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < per_image;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int vec = static_cast<int>(i % nvec);
    const float2 m = mr[(vec * 8) / cpg];
    const float mean = m.x, rstd = m.y;
    const uint4 u = ib[i];
    float v[8];
    float2 t;
    t = unpack_bf16(u.x); v[0] = t.x; v[1] = t.y;
    t = unpack_bf16(u.y); v[2] = t.x; v[3] = t.y;
    t = unpack_bf16(u.z); v[4] = t.x; v[5] = t.y;
    t = unpack_bf16(u.w); v[6] = t.x; v[7] = t.y;
    {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w) + vec * 2), w1 = __ldg(reinterpret_cast<const float4*>(w) + vec * 2 + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias) + vec * 2), b1 = __ldg(reinterpret_cast<const float4*>(bias) + vec * 2 + 1);
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (v[j] - mean) * rstd * wv[j] + bv[j];
    }
    uint4 o;
    o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]);
    o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
    ob[i] = o;
  }
}

// PreNorm (channel LayerNorm, gain only: module_util.py:77-97) fused with GroupNorm pass 1 for the SpatialTransformer
// levels (attention.py:251 normalises the PreNorm output): CTA (slab, image), one warp per pixel row.  The row is
// normalised, rounded to bf16 and stored (it is also the residual of proj_out), and the SAME rounded values feed the
// per-slab group sums - one launch and one read of the normalised tensor less than layernorm_rows + groupnorm_stats.
// Lane l holds the 8-channel vectors l, l + 32, ...; partial sums meet in shared memory, one thread per group adds them
// in a fixed order (bit-reproducible).
template <int VPL>
__global__ void __launch_bounds__(256) prenorm_gnstats_kernel(const __nv_bfloat16* __restrict__ in,
                                                              __nv_bfloat16* __restrict__ out, int hw, int c, int groups,
                                                              int slab, const float* __restrict__ g, float ln_eps,
                                                              float* __restrict__ stats) {
  __shared__ float2 part[8][32 * VPL];
  griddep_wait();
  griddep_launch();
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = c >> 3;
  const int p0 = blockIdx.x * slab, p1 = min(hw, p0 + slab);
  const float inv_c = 1.0f / c;
  float gs[VPL], gss[VPL];
  float gain[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    gs[i] = 0.f;
    gss[i] = 0.f;
    const int vi = lane + 32 * i;
    const float4 w0 = vi < nvec ? __ldg(reinterpret_cast<const float4*>(g) + vi * 2) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 w1 = vi < nvec ? __ldg(reinterpret_cast<const float4*>(g) + vi * 2 + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
    gain[i][0] = w0.x; gain[i][1] = w0.y; gain[i][2] = w0.z; gain[i][3] = w0.w;
    gain[i][4] = w1.x; gain[i][5] = w1.y; gain[i][6] = w1.z; gain[i][7] = w1.w;
  }
  for (int p = p0 + warp; p < p1; p += 8) {
    const int64_t row = static_cast<int64_t>(b) * hw + p;
    float v[VPL][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + 32 * i;
      uint4 u = make_uint4(0, 0, 0, 0);
      if (vi < nvec) u = __ldg(reinterpret_cast<const uint4*>(in + row * c) + vi);
      float2 t;
      t = unpack_bf16(u.x); v[i][0] = t.x; v[i][1] = t.y;
      t = unpack_bf16(u.y); v[i][2] = t.x; v[i][3] = t.y;
      t = unpack_bf16(u.z); v[i][4] = t.x; v[i][5] = t.y;
      t = unpack_bf16(u.w); v[i][6] = t.x; v[i][7] = t.y;
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[i][j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * inv_c;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (lane + 32 * i < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[i][j] - mean;
          ss = fmaf(d, d, ss);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss * inv_c + ln_eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = (v[i][j] - mean) * rstd * gain[i][j];
        uint4 o;
        o.x = pack_bf16(y[0], y[1]); o.y = pack_bf16(y[2], y[3]);
        o.z = pack_bf16(y[4], y[5]); o.w = pack_bf16(y[6], y[7]);
        reinterpret_cast<uint4*>(out + row * c)[vi] = o;
        // statistics of what GroupNorm will read: the bf16-rounded values
        const float2 a = unpack_bf16(o.x), bb = unpack_bf16(o.y), cc = unpack_bf16(o.z), d = unpack_bf16(o.w);
        gs[i] += (a.x + a.y) + (bb.x + bb.y) + (cc.x + cc.y) + (d.x + d.y);
        gss[i] += a.x * a.x + a.y * a.y + bb.x * bb.x + bb.y * bb.y + cc.x * cc.x + cc.y * cc.y + d.x * d.x + d.y * d.y;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) part[warp][lane + 32 * i] = make_float2(gs[i], gss[i]);
  __syncthreads();
  if (threadIdx.x < groups) {
    const int gi = threadIdx.x, vpg = (c / groups) >> 3;
    float s = 0.f, ss = 0.f;
    for (int w = 0; w < 8; ++w)
      for (int k = 0; k < vpg; ++k) {
        const float2 t = part[w][gi * vpg + k];
        s += t.x;
        ss += t.y;
      }
    reinterpret_cast<float2*>(stats)[(static_cast<int64_t>(b) * gridDim.x + blockIdx.x) * groups + gi] = make_float2(s, ss);
  }
}

// ------------------------------------------------------------------------------------------------ conditioning MLPs
// y[r] = W[r,:] . x + b[r] for r in [0, rows): ONE THREAD per output row (rows <= 2 * blockDim), 16-byte weight
// loads with four independent accumulators; x is broadcast from shared memory.  (A warp-per-row version with a
// shuffle reduction per row was latency-bound: 125 us for two 256x256 layers.)
__device__ __forceinline__ void block_linear(const float* __restrict__ W, const float* __restrict__ bias,
                                             const float* x_sh, float* y_sh, int rows, int k) {
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    const float4* wr = reinterpret_cast<const float4*>(W + static_cast<int64_t>(r) * k);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
    for (int i = 0; i < (k >> 2); ++i) {
      const float4 w4 = __ldg(wr + i);
      a0 = fmaf(w4.x, x_sh[4 * i], a0);
      a1 = fmaf(w4.y, x_sh[4 * i + 1], a1);
      a2 = fmaf(w4.z, x_sh[4 * i + 2], a2);
      a3 = fmaf(w4.w, x_sh[4 * i + 3], a3);
    }
    y_sh[r] = (a0 + a1) + (a2 + a3) + (bias ? __ldg(bias + r) : 0.f);
  }
}

// MODE 0: per step   - time MLP, + prompt embedding (precomputed per restoration), SiLU -> temb_out
// MODE 1: per restoration - text_mlp / softmax*prompt / prompt_mlp of the degradation context -> temb_out
template <int MODE>
__global__ void __launch_bounds__(256) time_embed_kernel(dac_embed_weights w, const float* __restrict__ time_ptr,
                                                         const float* __restrict__ text_ctx,
                                                         float* __restrict__ temb_out) {
  extern __shared__ float sh[];
  const int td = w.time_dim;
  float* a = sh;                 // [max(td, ctx, nf)]
  float* h = a + max(td, max(w.ctx_dim, w.nf));
  float* t = h + td;
  float* q = t + td;
  const int b = blockIdx.x;
  if (MODE == 0) {
  const float time = __ldg(time_ptr);
  // sinusoidal embedding (module_util.py:41-48): [sin | cos], freq_i = exp(-i ln(1e4)/(half-1))
  const int half = w.nf / 2;
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float f = expf(static_cast<float>(i) * -(logf(10000.0f) / static_cast<float>(half - 1)));
    const float arg = time * f;
    a[i] = sinf(arg);
    a[half + i] = cosf(arg);
  }
  __syncthreads();
  block_linear(w.time_w1, w.time_b1, a, h, td, w.nf);
  __syncthreads();
  for (int i = threadIdx.x; i < td; i += blockDim.x) h[i] = gelu_precise_f(h[i]);
  __syncthreads();
  block_linear(w.time_w2, w.time_b2, h, t, td, td);
  __syncthreads();
  // text_ctx holds the precomputed prompt embedding [B, td] here
  for (int i = threadIdx.x; i < td; i += blockDim.x) {
    const float v = t[i] + (text_ctx ? text_ctx[static_cast<int64_t>(b) * td + i] : 0.f);
    temb_out[static_cast<int64_t>(b) * td + i] = v / (1.0f + expf(-v));   // every ResBlock mlp starts with SiLU
  }
  return;
  }
  {
    for (int i = threadIdx.x; i < w.ctx_dim; i += blockDim.x) a[i] = text_ctx[static_cast<int64_t>(b) * w.ctx_dim + i];
    __syncthreads();
    block_linear(w.text_w1, w.text_b1, a, h, td, w.ctx_dim);
    __syncthreads();
    for (int i = threadIdx.x; i < td; i += blockDim.x) {
      const float v = h[i];
      h[i] = v / (1.0f + expf(-v));
    }
    __syncthreads();
    block_linear(w.text_w2, w.text_b2, h, q, td, td);
    __syncthreads();
    // softmax over the td features (arch.py:135), times the learned prompt
    __shared__ float red[2];
    if (threadIdx.x < 32) {
      float m = -INFINITY;
      for (int i = threadIdx.x; i < td; i += 32) m = fmaxf(m, q[i]);
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float s = 0.f;
      for (int i = threadIdx.x; i < td; i += 32) s += expf(q[i] - m);
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (threadIdx.x == 0) { red[0] = m; red[1] = s; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < td; i += blockDim.x) h[i] = expf(q[i] - red[0]) / red[1] * __ldg(w.prompt + i);
    __syncthreads();
    block_linear(w.prompt_w, w.prompt_b, h, q, td, td);
    __syncthreads();
    for (int i = threadIdx.x; i < td; i += blockDim.x) temb_out[static_cast<int64_t>(b) * td + i] = q[i];
  }
}

__global__ void __launch_bounds__(256) film_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                                   const float* __restrict__ s, float* __restrict__ film, int F, int k,
                                                   int B) {
  // silu(t_emb) of every image staged in shared memory; one warp per output feature f (grid-stride), its weight
  // row held in registers (k <= 256 -> 8 per lane), looping over the images.
  extern __shared__ float s_sh[];
  for (int i = threadIdx.x; i < B * k; i += blockDim.x) s_sh[i] = s[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; f < F; f += nw) {
    float wr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      wr[j] = (lane + 32 * j < k) ? __ldg(W + static_cast<int64_t>(f) * k + lane + 32 * j) : 0.f;
    const float bf = __ldg(bias + f);
    for (int b = 0; b < B; ++b) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (lane + 32 * j < k) acc = fmaf(wr[j], s_sh[b * k + lane + 32 * j], acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) film[static_cast<int64_t>(b) * F + f] = acc + bf;
    }
  }
}

__global__ void __launch_bounds__(256) two_linear_kernel(const float* __restrict__ x, int in,
                                                         const float* __restrict__ w1, int mid,
                                                         const float* __restrict__ w2, const float* __restrict__ b2,
                                                         int out, float* __restrict__ y) {
  extern __shared__ float sh[];
  float* xs = sh;
  float* ms = sh + in;
  float* ys = ms + mid;
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < in; i += blockDim.x) xs[i] = x[static_cast<int64_t>(b) * in + i];
  __syncthreads();
  block_linear(w1, nullptr, xs, ms, mid, in);
  __syncthreads();
  block_linear(w2, b2, ms, ys, out, mid);
  __syncthreads();
  for (int i = threadIdx.x; i < out; i += blockDim.x) y[static_cast<int64_t>(b) * out + i] = ys[i];
}

}  // namespace dac

using namespace dac;

extern "C" int dac_sde_step(int mode, const float* x, const float* mu, const float* net, const float* eps, float* out,
                            int64_t n, const float* host_coef, dac_stream_t stream) {
  if (!x || !mu || !net || !out || !host_coef || (mode != 2 && !eps)) return set_error(-1, "dac_sde_step: null argument");
  if (n <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(mu) | reinterpret_cast<uintptr_t>(net) |
       reinterpret_cast<uintptr_t>(eps) | reinterpret_cast<uintptr_t>(out)) & 15)
    return set_error(-2, "dac_sde_step: pointers must be 16-byte aligned");
  SdeCoef k;
  for (int i = 0; i < 8; ++i) k.c[i] = host_coef[i];
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = elementwise_grid(n / 4 + 1);
  if (mode == 0) sde_step_kernel<0><<<grid, 256, 0, s>>>(x, mu, net, eps, out, n, k);
  else if (mode == 1) sde_step_kernel<1><<<grid, 256, 0, s>>>(x, mu, net, eps, out, n, k);
  else if (mode == 2) sde_step_kernel<2><<<grid, 256, 0, s>>>(x, mu, net, eps, out, n, k);
  else return set_error(-2, "dac_sde_step: unknown mode %d", mode);
  return check_launch("sde_step_kernel");
}

extern "C" int dac_loop_tick(int64_t* state, const float* t_table, const float* coef_table, float* t_dev,
                             float* coef_dev, dac_stream_t stream) {
  if (!state || !t_table || !coef_table || !t_dev || !coef_dev) return set_error(-1, "dac_loop_tick: null argument");
  loop_tick_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<long long*>(state), t_table, coef_table,
                                                                    t_dev, coef_dev);
  return check_launch("loop_tick_kernel");
}

extern "C" int dac_sde_step_dev(int mode, const float* x, const float* mu, const float* net, float* out, int64_t n,
                                const float* coef_dev, const int64_t* state, dac_stream_t stream) {
  if (!x || !mu || !net || !out || !coef_dev || !state) return set_error(-1, "dac_sde_step_dev: null argument");
  if (n <= 0 || (n & 3)) return set_error(-2, "dac_sde_step_dev: n must be a positive multiple of 4");
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(mu) | reinterpret_cast<uintptr_t>(net) |
       reinterpret_cast<uintptr_t>(out)) & 15)
    return set_error(-2, "dac_sde_step_dev: pointers must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = elementwise_grid(n / 4 + 1);
  const long long* st = reinterpret_cast<const long long*>(state);
  if (mode == 0) launch_k(sde_step_dev_kernel<0>, dim3(grid), dim3(256), 0, s, x, mu, net, out, n, coef_dev, st);
  else if (mode == 1) launch_k(sde_step_dev_kernel<1>, dim3(grid), dim3(256), 0, s, x, mu, net, out, n, coef_dev, st);
  else if (mode == 2) launch_k(sde_step_dev_kernel<2>, dim3(grid), dim3(256), 0, s, x, mu, net, out, n, coef_dev, st);
  else return set_error(-2, "dac_sde_step_dev: unknown mode %d", mode);
  return check_launch("sde_step_dev_kernel");
}

extern "C" int dac_noise_state(const float* x, const float* eps, float* out, int64_t n, float max_sigma,
                               dac_stream_t stream) {
  if (!x || !eps || !out) return set_error(-1, "dac_noise_state: null argument");
  if (n <= 0) return 0;
  noise_state_kernel<<<elementwise_grid(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, eps, out, n, max_sigma);
  return check_launch("noise_state_kernel");
}

extern "C" int dac_unet_stem_input(const float* xt, const float* cond, void* out, int B, int H, int W, int Hp, int Wp,
                                   int pair, dac_stream_t stream) {
  if (!xt || !cond || !out) return set_error(-1, "dac_unet_stem_input: null argument");
  if (Hp < H || Wp < W || Hp - H >= H || Wp - W >= W) return set_error(-2, "dac_unet_stem_input: bad padding");
  if (Hp > 65535 || B > 65535) return set_error(-2, "dac_unet_stem_input: Hp and B must be < 65536");
  if (Wp > 2000) return set_error(-2, "dac_unet_stem_input: padded width must be <= 2000 (one row in 48 KB of shared memory)");
  if (pair && (Wp & 1)) return set_error(-2, "dac_unet_stem_input: the pixel-pair packing needs an even padded width");
  if (pair)
    stem_input_kernel<true><<<dim3(Hp, B), 256, sizeof(float) * 6 * (Wp + 6), static_cast<cudaStream_t>(stream)>>>(
        xt, cond, static_cast<__nv_bfloat16*>(out), B, H, W, Hp, Wp);
  else
    stem_input_kernel<false><<<dim3(Hp, B), 256, sizeof(float) * 6 * (Wp + 6), static_cast<cudaStream_t>(stream)>>>(
        xt, cond, static_cast<__nv_bfloat16*>(out), B, H, W, Hp, Wp);
  return check_launch("stem_input_kernel");
}

extern "C" int dac_layernorm_rows(const void* in, int32_t ld_in, void* out, int32_t ld_out, int64_t rows, int32_t c,
                                  const float* w, const float* b, float eps, dac_stream_t stream) {
  return dac_layernorm_rows_res(in, ld_in, nullptr, 0, out, ld_out, rows, c, w, b, eps, stream);
}

extern "C" int dac_layernorm_rows_res(const void* in, int32_t ld_in, const void* res, int32_t ld_res, void* out,
                                      int32_t ld_out, int64_t rows, int32_t c, const float* w, const float* b, float eps,
                                      dac_stream_t stream) {
  if (!in || !out) return set_error(-1, "dac_layernorm_rows: null argument");
  if ((c & 7) || (ld_in & 7) || (ld_out & 7) || (ld_res & 7))
    return set_error(-2, "dac_layernorm_rows: c and pitches must be multiples of 8");
  const __nv_bfloat16* rp = static_cast<const __nv_bfloat16*>(res);
  if (rows <= 0) return 0;
  if (c > 1024) return set_error(-2, "dac_layernorm_rows: c must be <= 1024");
  if ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(b)) & 15)
    return set_error(-2, "dac_layernorm_rows: w and b must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* ip = static_cast<const __nv_bfloat16*>(in);
  __nv_bfloat16* op = static_cast<__nv_bfloat16*>(out);
  const int nvec = c >> 3;
  const int lpr = nvec <= 8 ? 8 : (nvec <= 16 ? 16 : 32);
  const int64_t warps = ceil_div(rows, 32 / lpr);
  int64_t blocks = ceil_div(warps, 8);
  // persistent warps (the parameters are loaded once per warp): as many CTAs as are resident at once - 48 / 76 / 128
  // registers per thread for 1 / 2 / 4 vectors per lane
  const int resident = 148 * (nvec <= 32 ? 4 : (nvec <= 64 ? 3 : 2));
  if (blocks > resident) blocks = resident;
  const int grid = static_cast<int>(blocks);
  const dim3 g(grid), t(256);
  const long long rows_ll = rows;
  if (lpr == 8) launch_k(layernorm_rows_kernel<8, 1>, g, t, 0, st, ip, ld_in, op, ld_out, rows_ll, c, w, b, eps, rp, ld_res);
  else if (lpr == 16) launch_k(layernorm_rows_kernel<16, 1>, g, t, 0, st, ip, ld_in, op, ld_out, rows_ll, c, w, b, eps, rp, ld_res);
  else if (nvec <= 32) launch_k(layernorm_rows_kernel<32, 1>, g, t, 0, st, ip, ld_in, op, ld_out, rows_ll, c, w, b, eps, rp, ld_res);
  else if (nvec <= 64) launch_k(layernorm_rows_kernel<32, 2>, g, t, 0, st, ip, ld_in, op, ld_out, rows_ll, c, w, b, eps, rp, ld_res);
  else launch_k(layernorm_rows_kernel<32, 4>, g, t, 0, st, ip, ld_in, op, ld_out, rows_ll, c, w, b, eps, rp, ld_res);
  return check_launch("layernorm_rows_kernel");
}

extern "C" int dac_layernorm_rows_f32(const float* in, int32_t ld_in, void* out, int32_t ld_out, int64_t rows,
                                      int32_t c, const float* w, const float* b, float eps, dac_stream_t stream) {
  if (!in || !out) return set_error(-1, "dac_layernorm_rows_f32: null argument");
  if ((c & 3) || (ld_in & 3) || (ld_out & 3)) return set_error(-2, "dac_layernorm_rows_f32: c and pitches must be multiples of 4");
  if (rows <= 0) return 0;
  if (c > 1024) return set_error(-2, "dac_layernorm_rows_f32: c must be <= 1024");
  if ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(in)) & 15)
    return set_error(-2, "dac_layernorm_rows_f32: in, w and b must be 16-byte aligned");
  // persistent warps (two 256-thread CTAs per SM): weight / bias stay in registers across a warp's rows
  const int64_t blocks = ceil_div(rows, 8);
  const int grid = static_cast<int>(blocks > 148 * 2 ? 148 * 2 : blocks);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* op = static_cast<__nv_bfloat16*>(out);
  const int nv = (c / 4 + 31) / 32;
  if (nv <= 2) layernorm_rows_f32_kernel<2><<<grid, 256, 0, st>>>(in, ld_in, op, ld_out, rows, c, w, b, eps);
  else if (nv <= 4) layernorm_rows_f32_kernel<4><<<grid, 256, 0, st>>>(in, ld_in, op, ld_out, rows, c, w, b, eps);
  else if (nv <= 6) layernorm_rows_f32_kernel<6><<<grid, 256, 0, st>>>(in, ld_in, op, ld_out, rows, c, w, b, eps);
  else layernorm_rows_f32_kernel<8><<<grid, 256, 0, st>>>(in, ld_in, op, ld_out, rows, c, w, b, eps);
  return check_launch("layernorm_rows_f32_kernel");
}

extern "C" int dac_groupnorm_nhwc(const void* in, void* out, int32_t B, int32_t hw, int32_t c, int32_t groups,
                                  const float* w, const float* b, float eps, float* stats, dac_stream_t stream) {
  if (!in || !out || !w || !b || !stats) return set_error(-1, "dac_groupnorm_nhwc: null argument");
  if (c % groups || (c / groups) % 8 || c / 8 > 256) return set_error(-2, "dac_groupnorm_nhwc: need 8 | c/groups, c <= 2048");
  if ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(b)) & 15)
    return set_error(-2, "dac_groupnorm_nhwc: w and b must be 16-byte aligned");
  if (groups > 256 || B > 65535) return set_error(-2, "dac_groupnorm_nhwc: groups <= 256, B <= 65535");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int slab = static_cast<int>(ceil_div(hw, kGnSlabs));
  launch_k(groupnorm_stats_kernel, dim3(kGnSlabs, B), dim3(256), 0, s, static_cast<const __nv_bfloat16*>(in), hw, c, groups,
           slab, stats);
  int rc = check_launch("groupnorm_stats_kernel");
  if (rc) return rc;
  const int64_t per_image = static_cast<int64_t>(hw) * (c / 8);
  int64_t chunks = ceil_div(per_image, 256 * 4);
  const int64_t cap = ceil_div(148 * 16, B);
  if (chunks > cap) chunks = cap;
  launch_k(groupnorm_apply_kernel, dim3(static_cast<unsigned>(chunks < 1 ? 1 : chunks), B), dim3(256), 0, s,
           static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), hw, c, groups, kGnSlabs, w, b, eps,
           static_cast<const float*>(stats));
  return check_launch("groupnorm_apply_kernel");
}

extern "C" int dac_prenorm_groupnorm_nhwc(const void* in, void* normed, void* out, int32_t B, int32_t hw, int32_t c,
                                          int32_t groups, const float* pre_g, float pre_eps, const float* w, const float* b,
                                          float eps, float* stats, dac_stream_t stream) {
  if (!in || !normed || !out || !pre_g || !w || !b || !stats) return set_error(-1, "dac_prenorm_groupnorm_nhwc: null argument");
  if (c % groups || (c / groups) % 8 || c > 1024 || (c & 7))
    return set_error(-2, "dac_prenorm_groupnorm_nhwc: need 8 | c/groups, c <= 1024");
  if ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(pre_g)) & 15)
    return set_error(-2, "dac_prenorm_groupnorm_nhwc: pre_g, w and b must be 16-byte aligned");
  if (groups > 256 || B > 65535) return set_error(-2, "dac_prenorm_groupnorm_nhwc: groups <= 256, B <= 65535");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int slab = static_cast<int>(ceil_div(hw, kGnSlabs));
  const __nv_bfloat16* ip = static_cast<const __nv_bfloat16*>(in);
  __nv_bfloat16* np = static_cast<__nv_bfloat16*>(normed);
  const int vpl = (c / 8 + 31) / 32;
  const dim3 g1(kGnSlabs, B), t(256);
  if (vpl == 1) launch_k(prenorm_gnstats_kernel<1>, g1, t, 0, s, ip, np, hw, c, groups, slab, pre_g, pre_eps, stats);
  else if (vpl == 2) launch_k(prenorm_gnstats_kernel<2>, g1, t, 0, s, ip, np, hw, c, groups, slab, pre_g, pre_eps, stats);
  else launch_k(prenorm_gnstats_kernel<4>, g1, t, 0, s, ip, np, hw, c, groups, slab, pre_g, pre_eps, stats);
  int rc = check_launch("prenorm_gnstats_kernel");
  if (rc) return rc;
  const int64_t per_image = static_cast<int64_t>(hw) * (c / 8);
  int64_t chunks = ceil_div(per_image, 256 * 4);
  const int64_t cap = ceil_div(148 * 16, B);
  if (chunks > cap) chunks = cap;
  launch_k(groupnorm_apply_kernel, dim3(static_cast<unsigned>(chunks < 1 ? 1 : chunks), B), dim3(256), 0, s,
           static_cast<const __nv_bfloat16*>(np), static_cast<__nv_bfloat16*>(out), hw, c, groups, kGnSlabs, w, b, eps,
           static_cast<const float*>(stats));
  return check_launch("groupnorm_apply_kernel");
}

static size_t embed_smem(const dac_embed_weights* w) {
  int amax = w->time_dim > w->ctx_dim ? w->time_dim : w->ctx_dim;
  if (w->nf > amax) amax = w->nf;
  return sizeof(float) * (amax + 3 * w->time_dim);
}

extern "C" int dac_prompt_embed(const dac_embed_weights* w, const float* text_ctx, int32_t B, float* prompt_emb,
                                dac_stream_t stream) {
  if (!w || !text_ctx || !prompt_emb || !w->text_w1) return set_error(-1, "dac_prompt_embed: null argument");
  if (w->time_dim > 256 || w->time_dim % 32) return set_error(-2, "dac_prompt_embed: time_dim must be <= 256");
  time_embed_kernel<1><<<B, 256, embed_smem(w), static_cast<cudaStream_t>(stream)>>>(*w, nullptr, text_ctx, prompt_emb);
  return check_launch("time_embed_kernel<prompt>");
}

extern "C" int dac_time_film(const dac_embed_weights* w, const float* time, const float* prompt_emb, int32_t B,
                             float* temb_scratch, float* film, dac_stream_t stream) {
  if (!w || !time || !temb_scratch || !film) return set_error(-1, "dac_time_film: null argument");
  if (w->time_dim > 256 || w->time_dim % 32) return set_error(-2, "dac_time_film: time_dim must be <= 256");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int td = w->time_dim;
  time_embed_kernel<0><<<B, 256, embed_smem(w), s>>>(*w, time, prompt_emb, temb_scratch);
  int rc = check_launch("time_embed_kernel");
  if (rc) return rc;
  const size_t fsh = sizeof(float) * B * td;
  if (fsh > 160 * 1024) return set_error(-2, "dac_time_film: batch too large for the FiLM kernel's shared memory");
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(film_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    attr_done = true;
  }
  int grid = static_cast<int>(ceil_div(w->F, 8 * 4));
  if (grid > 296) grid = 296;
  film_kernel<<<grid, 256, fsh, s>>>(w->film_w, w->film_b, temb_scratch, film, w->F, td, B);
  return check_launch("film_kernel");
}

extern "C" int dac_two_linear(const float* x, int32_t B, int32_t in, const float* w1, int32_t mid, const float* w2,
                              const float* b2, int32_t out, float* y, dac_stream_t stream) {
  if (!x || !w1 || !w2 || !y) return set_error(-1, "dac_two_linear: null argument");
  const size_t sh = sizeof(float) * (in + mid + out);
  if (sh > 48 * 1024) return set_error(-2, "dac_two_linear: dims too large");
  two_linear_kernel<<<B, 256, sh, static_cast<cudaStream_t>(stream)>>>(x, in, w1, mid, w2, b2, out, y);
  return check_launch("two_linear_kernel");
}
