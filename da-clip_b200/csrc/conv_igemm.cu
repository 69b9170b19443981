// Implicit-GEMM convolution / linear layer on tcgen05 tensor cores (sm_100a).
//
//   D[pixel, cout] = sum_{tap, cin} X[pixel + tap, cin] * W[tap][cout][cin]
//
// M = 128 output pixels (a tile_h x tile_w box of one image), N = block_n output channels,
// K walks (tap, source, 64-channel chunk).  One persistent CTA per SM, three warp roles:
//   warp 0      TMA producer: per K step one 4-D box load of the (shifted) activation tile - image borders
//               are zero-filled by TMA, which IS the conv padding - plus one 3-D box load of the weight slab,
//               both 128B-swizzled, into an smem ring guarded by full/empty mbarriers;
//   warp 1      allocates TMEM, one lane issues tcgen05.mma (M128 x N x K16, bf16 -> fp32 in TMEM) and
//               tcgen05.commit's the smem slot back to the producer / the accumulator to the epilogue;
//   warps 2-5   epilogue: tcgen05.ld the accumulator (one pixel per thread, all channels thread-local),
//               apply bias / FiLM / SiLU|GELU / GEGLU / channel-LayerNorm / q-softmax / residual in fp32,
//               store NHWC bf16 (or fp32 NCHW for final_conv).  Two TMEM accumulator stages let the
//               epilogue of tile i overlap the MMAs of tile i+1.
// Replaces nn.Conv2d / nn.Linear + the pointwise ops around them in
//   module_util.py:100-153,157-185 and attention.py:37-64,152-261 of the reference.
#include <cuda.h>
#include <cuda_runtime.h>
#include <mutex>
#include <new>
#include <stdio.h>
#include <string.h>

#include "../../include/dac_b200.h"
#include "common.h"
#include "ptx.cuh"

namespace dac {

constexpr int kTileM = 128;
constexpr int kChunkK = 64;  // bf16 per K step = one 128 B swizzle row
constexpr int kThreads = 192;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccStride = 256;
constexpr uint32_t kABytes = kTileM * kChunkK * 2;  // 16 KB
constexpr int kMaxStages = 8;

struct ConvKParams {
  int B, OH, OW, stride;
  int tile_h, tile_w, tiles_x, tiles_y, m_tiles;
  int n_tiles, block_n, ngroups, ntaps;
  int chunks0, chunks1, c0;
  int per_image_w;
  int stages;
  uint32_t b_bytes;
  int8_t tap_dy[4][16];
  int8_t tap_dx[4][16];
  // epilogue
  int epi, act, cout;
  const float* bias;
  const float* bias_img;
  const float* film;
  int film_ld, film_off;
  const float* ln_g;
  float ln_eps;
  const __nv_bfloat16* res;
  int res_ld;
  const __nv_bfloat16* res2;
  int res2_ld;
  __nv_bfloat16* out;
  int out_ld, out_coff;
  const float* res_f32;
  int res_f32_ld;
  float* out_f32;
  int out_f32_ld;
  int out_scale, OHf, OWf;
  int8_t out_oy[4], out_ox[4];
  float* out_nchw;
  int nchw_c, nchw_h, nchw_w;
};

struct TileCoord {
  int g, nt, n, y0, x0;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvKParams& p, int tile) {
  TileCoord t;
  t.nt = tile % p.n_tiles;
  int rest = tile / p.n_tiles;
  int mt = rest % p.m_tiles;
  t.g = rest / p.m_tiles;
  int tx = mt % p.tiles_x;
  int r2 = mt / p.tiles_x;
  int ty = r2 % p.tiles_y;
  t.n = r2 / p.tiles_y;
  t.y0 = ty * p.tile_h;
  t.x0 = tx * p.tile_w;
  return t;
}

template <int CW>
__device__ __forceinline__ void tmem_ld_cw(uint32_t taddr, uint32_t (&r)[CW]);
template <>
__device__ __forceinline__ void tmem_ld_cw<32>(uint32_t taddr, uint32_t (&r)[32]) {
  tmem_ld32(taddr, r);
}
template <>
__device__ __forceinline__ void tmem_ld_cw<16>(uint32_t taddr, uint32_t (&r)[16]) {
  tmem_ld16(taddr, r);
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == DAC_ACT_SILU) return silu_f(v);
  if (act == DAC_ACT_GELU) return gelu_f(v);
  return v;
}

// Store CW consecutive channels of one pixel.
template <int CW>
__device__ __forceinline__ void store_bf16_chunk(__nv_bfloat16* dst, const float (&v)[CW], int nvalid) {
  if (nvalid >= CW) {
#pragma unroll
    for (int q = 0; q < CW / 8; ++q) {
      uint4 u;
      u.x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
      u.y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
      u.z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
      u.w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
      reinterpret_cast<uint4*>(dst)[q] = u;
    }
  } else {
    for (int j = 0; j < nvalid; ++j) dst[j] = __float2bfloat16(v[j]);
  }
}

template <int CW>
__device__ __forceinline__ void add_residual(const __nv_bfloat16* src, float (&v)[CW], int nvalid) {
  if (nvalid >= CW) {
#pragma unroll
    for (int q = 0; q < CW / 8; ++q) {
      uint4 u = __ldg(reinterpret_cast<const uint4*>(src) + q);
      float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
      v[q * 8 + 0] += a.x; v[q * 8 + 1] += a.y; v[q * 8 + 2] += b.x; v[q * 8 + 3] += b.y;
      v[q * 8 + 4] += c.x; v[q * 8 + 5] += c.y; v[q * 8 + 6] += d.x; v[q * 8 + 7] += d.y;
    }
  } else {
    for (int j = 0; j < nvalid; ++j) v[j] += __bfloat162float(src[j]);
  }
}

// Epilogue for one 128 x block_n accumulator tile; executed by the 4 epilogue warps.
template <int CW>
__device__ __forceinline__ void epilogue_tile(const ConvKParams& p, const TileCoord& t, uint32_t tmem_acc,
                                              int row) {
  const int ty = row / p.tile_w, tx = row - ty * p.tile_w;
  const int y = t.y0 + ty, x = t.x0 + tx;
  const bool valid = (y < p.OH) && (x < p.OW);
  const int Y = y * p.out_scale + p.out_oy[t.g], X = x * p.out_scale + p.out_ox[t.g];
  const long long opix = (static_cast<long long>(t.n) * p.OHf + Y) * p.OWf + X;
  const int n = t.n;
  uint32_t r[CW];
  float v[CW];

  if (p.epi == DAC_EPI_LN) {
    // channel LayerNorm over the whole (single) N tile: 3 sweeps over TMEM, all thread-local.
    const int C = p.cout;
    float sum = 0.f;
    for (int c = 0; c < C; c += CW) {
      tmem_ld_cw<CW>(tmem_acc + c, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < CW; ++j) sum += __uint_as_float(r[j]) + (p.bias ? __ldg(p.bias + c + j) : 0.f);
    }
    const float mean = sum / C;
    float ss = 0.f;
    for (int c = 0; c < C; c += CW) {
      tmem_ld_cw<CW>(tmem_acc + c, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < CW; ++j) {
        float d = __uint_as_float(r[j]) + (p.bias ? __ldg(p.bias + c + j) : 0.f) - mean;
        ss += d * d;
      }
    }
    const float rstd = rsqrtf(ss / C + p.ln_eps);
    for (int c = 0; c < C; c += CW) {
      tmem_ld_cw<CW>(tmem_acc + c, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < CW; ++j) {
        float a = __uint_as_float(r[j]) + (p.bias ? __ldg(p.bias + c + j) : 0.f);
        v[j] = (a - mean) * rstd * __ldg(p.ln_g + c + j);
      }
      if (valid) {
        if (p.res) add_residual<CW>(p.res + opix * p.res_ld + c, v, CW);
        store_bf16_chunk<CW>(p.out + opix * p.out_ld + p.out_coff + c, v, CW);
      }
    }
    return;
  }

  if (p.epi == DAC_EPI_GEGLU) {
    const int half = p.block_n >> 1;
    for (int c = 0; c < half; c += CW) {
      uint32_t rg[CW];
      tmem_ld_cw<CW>(tmem_acc + c, r);
      tmem_ld_cw<CW>(tmem_acc + half + c, rg);
      tmem_ld_wait();
      const int col = t.nt * p.block_n + c;  // column in the (permuted) weight/bias row order
#pragma unroll
      for (int j = 0; j < CW; ++j) {
        float a = __uint_as_float(r[j]) + __ldg(p.bias + col + j);
        float g = __uint_as_float(rg[j]) + __ldg(p.bias + col + half + j);
        v[j] = a * gelu_f(g);
      }
      if (valid) store_bf16_chunk<CW>(p.out + opix * p.out_ld + p.out_coff + t.nt * half + c, v, CW);
    }
    return;
  }

  // PLAIN and QKV
  for (int c = 0; c < p.block_n; c += CW) {
    const int ch = t.nt * p.block_n + c;
    if (ch >= p.cout) break;  // warp-uniform
    tmem_ld_cw<CW>(tmem_acc + c, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < CW; ++j) v[j] = __uint_as_float(r[j]);
    const int nvalid = min(CW, p.cout - ch);
    if (p.epi == DAC_EPI_QKV) {
      if (t.nt == 0) {  // q: softmax over the 32 channels of one head, times dim_head^-0.5
        float m = v[0];
#pragma unroll
        for (int j = 1; j < CW; ++j) m = fmaxf(m, v[j]);
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < CW; ++j) {
          v[j] = __expf(v[j] - m);
          s += v[j];
        }
        const float inv = 0.17677669529663687f / s;
#pragma unroll
        for (int j = 0; j < CW; ++j) v[j] *= inv;
      }
    } else {
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < CW; ++j)
          if (j < nvalid) v[j] += __ldg(p.bias + ch + j);
      }
      if (p.bias_img) {
        const float* bi = p.bias_img + static_cast<long long>(n) * p.cout + ch;
#pragma unroll
        for (int j = 0; j < CW; ++j)
          if (j < nvalid) v[j] += __ldg(bi + j);
      }
      if (p.film) {
        const float* sc = p.film + static_cast<long long>(n) * p.film_ld + p.film_off + ch;
#pragma unroll
        for (int j = 0; j < CW; ++j)
          if (j < nvalid) v[j] = v[j] * (__ldg(sc + j) + 1.0f) + __ldg(sc + p.cout + j);
      }
      if (p.act != DAC_ACT_NONE) {
#pragma unroll
        for (int j = 0; j < CW; ++j) v[j] = apply_act(v[j], p.act);
      }
    }
    if (valid) {
      if (p.out_nchw) {
        // fp32 planar output (final_conv): cropped to the un-padded image.
        if (Y < p.nchw_h && X < p.nchw_w) {
          for (int j = 0; j < nvalid && ch + j < p.nchw_c; ++j)
            p.out_nchw[((static_cast<long long>(n) * p.nchw_c + ch + j) * p.nchw_h + Y) * p.nchw_w + X] = v[j];
        }
      } else {
        if (p.res_f32) {
          const float* rp = p.res_f32 + opix * p.res_f32_ld + ch;
#pragma unroll
          for (int j = 0; j < CW; ++j)
            if (j < nvalid) v[j] += __ldg(rp + j);
        }
        if (p.res) add_residual<CW>(p.res + opix * p.res_ld + ch, v, nvalid);
        if (p.res2) add_residual<CW>(p.res2 + opix * p.res2_ld + ch, v, nvalid);
        if (p.out_f32) {
          float* op = p.out_f32 + opix * p.out_f32_ld + ch;
          if (nvalid >= CW) {
#pragma unroll
            for (int q = 0; q < CW / 4; ++q)
              reinterpret_cast<float4*>(op)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          } else {
            for (int j = 0; j < nvalid; ++j) op[j] = v[j];
          }
        }
        if (p.out) store_bf16_chunk<CW>(p.out + opix * p.out_ld + p.out_coff + ch, v, nvalid);
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                  const __grid_constant__ CUtensorMap mapW, const __grid_constant__ ConvKParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024 B alignment is required by the 128B swizzle atoms (TMA write and UMMA read agree on address bits).
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t stage_bytes = kABytes + p.b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.stages) * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kMaxStages;
  uint64_t* tmem_full = bars + 2 * kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.ngroups * p.m_tiles * p.n_tiles;
  const int k_steps = p.ntaps * (p.chunks0 + p.chunks1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapW);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
#ifdef DAC_DEBUG
  if (threadIdx.x == 0 && blockIdx.x == 0)
    printf("[conv] tiles=%d k_steps=%d stages=%d block_n=%d tmem_base=%08x smem=%p OH=%d OW=%d out=%p cout=%d epi=%d\n",
           total_tiles, k_steps, p.stages, p.block_n, tmem_base, smem, p.OH, p.OW, p.out, p.cout, p.epi);
#endif

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, tile);
        const int xin = t.x0 * p.stride, yin = t.y0 * p.stride;
        const int zbase = (p.per_image_w ? t.n * p.ngroups * p.ntaps : 0) + t.g * p.ntaps;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const int dy = p.tap_dy[t.g][tap], dx = p.tap_dx[t.g][tap];
          for (int ck = 0; ck < p.chunks0 + p.chunks1; ++ck) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = smem + static_cast<size_t>(stage) * stage_bytes;
            mbar_arrive_expect_tx(&full[stage], stage_bytes);
            if (ck < p.chunks0)
              tma_load_4d(sa, &mapA0, &full[stage], ck * kChunkK, xin + dx, yin + dy, t.n);
            else
              tma_load_4d(sa, &mapA1, &full[stage], (ck - p.chunks0) * kChunkK, xin + dx, yin + dy, t.n);
            tma_load_3d(sa + kABytes, &mapW, &full[stage], ck * kChunkK, t.nt * p.block_n, zbase + tap);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc_bf16(kTileM, p.block_n);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kAccStride;
      for (int ks = 0; ks < k_steps; ++ks) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(smem + static_cast<size_t>(stage) * stage_bytes);
          const uint64_t adesc = make_sw128_desc(a_addr);
          const uint64_t bdesc = make_sw128_desc(a_addr + kABytes);
#pragma unroll
          for (int k = 0; k < kChunkK / 16; ++k)
            umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (ks | k) != 0 ? 1u : 0u);
          umma_commit(&empty[stage]);
          if (ks == k_steps - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue warps (2..5) =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
#ifdef DAC_DEBUG
      if (threadIdx.x == 64 && blockIdx.x == 0) printf("[conv] epilogue tile %d acc %d\n", tile, acc);
#endif
      const uint32_t tmem_acc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kAccStride;
      if (p.block_n & 31)
        epilogue_tile<16>(p, t, tmem_acc, row);
      else
        epilogue_tile<32>(p, t, tmem_acc, row);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    // No link-time dependency on libcuda: the library must load on hosts without a driver.
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(f);
  });
  return fn;
}

}  // namespace dac

struct dac_conv_plan {
  CUtensorMap mapA0, mapA1, mapW;
  dac::ConvKParams kp;
  int grid;
  int smem;
  int tiles;
};

using namespace dac;

static int encode_act_map(CUtensorMap* m, const void* ptr, int c, int ld, int W, int H, int B, int tile_w,
                          int tile_h, int stride) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return set_error(-10, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {(cuuint32_t)kChunkK, (cuuint32_t)(tile_w * stride), (cuuint32_t)(tile_h * stride), 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(-11, "cuTensorMapEncodeTiled(activation) failed: CUresult %d", (int)r);
  return 0;
}

extern "C" int dac_conv_create(const dac_conv_desc* d, dac_conv_t* out) {
  if (!d || !out) return set_error(-1, "dac_conv_create: null argument");
  *out = nullptr;
  if (d->tile_h * d->tile_w != kTileM || d->tile_w < 8 || (d->tile_w & 7))
    return set_error(-2, "dac_conv_create: tile %dx%d must have 128 pixels, width multiple of 8", d->tile_h, d->tile_w);
  if (d->c0 <= 0 || d->c0 % kChunkK || d->c1 % kChunkK || d->ld0 % 8 || (d->c1 && d->ld1 % 8))
    return set_error(-2, "dac_conv_create: channels (%d,%d) must be multiples of 64, pitches of 8", d->c0, d->c1);
  if (d->block_n < 16 || d->block_n > 256 || d->block_n % 16 || d->cout_pad % d->block_n || d->cout > d->cout_pad)
    return set_error(-2, "dac_conv_create: bad block_n %d / cout %d / cout_pad %d", d->block_n, d->cout, d->cout_pad);
  if (d->ntaps < 1 || d->ntaps > 16 || (d->ngroups != 1 && d->ngroups != 4) || (d->stride != 1 && d->stride != 2))
    return set_error(-2, "dac_conv_create: bad ntaps/ngroups/stride");
  if (d->tile_w * d->stride > 256 || d->tile_h * d->stride > 256)
    return set_error(-2, "dac_conv_create: TMA box exceeds 256");
  if ((reinterpret_cast<uintptr_t>(d->src0) | reinterpret_cast<uintptr_t>(d->src1) |
       reinterpret_cast<uintptr_t>(d->weight) | reinterpret_cast<uintptr_t>(d->out) |
       reinterpret_cast<uintptr_t>(d->res) | reinterpret_cast<uintptr_t>(d->res2)) & 15)
    return set_error(-2, "dac_conv_create: pointers must be 16-byte aligned");
  if (!d->out && !d->out_nchw && !d->out_f32) return set_error(-2, "dac_conv_create: no output");
  if ((d->out_f32 && (d->out_f32_ld & 3)) || (d->res_f32 && (d->res_f32_ld & 3)) ||
      ((reinterpret_cast<uintptr_t>(d->out_f32) | reinterpret_cast<uintptr_t>(d->res_f32)) & 15))
    return set_error(-2, "dac_conv_create: fp32 stream pointers/pitches must be 16-byte aligned");
  if ((d->out_f32 || d->res_f32) && d->epi != DAC_EPI_PLAIN)
    return set_error(-2, "dac_conv_create: fp32 stream only with the PLAIN epilogue");
  if (d->epi == DAC_EPI_LN && (d->cout != d->block_n || d->cout_pad != d->cout || !d->ln_g || (d->cout & 31)))
    return set_error(-2, "dac_conv_create: LN epilogue needs a single N tile with cout %% 32 == 0");
  if (d->epi == DAC_EPI_QKV && (d->block_n != 128 || d->cout != 384))
    return set_error(-2, "dac_conv_create: QKV epilogue needs block_n 128, cout 384");
  if (d->epi == DAC_EPI_GEGLU && (!d->bias || (d->block_n & 63) || d->cout != d->cout_pad))
    return set_error(-2, "dac_conv_create: GEGLU epilogue needs bias and block_n %% 64 == 0");
  if (d->out && ((d->out_ld | d->out_coff) & 7)) return set_error(-2, "dac_conv_create: out_ld/out_coff %% 8");
  if ((d->res && (d->res_ld & 7)) || (d->res2 && (d->res2_ld & 7) ))
    return set_error(-2, "dac_conv_create: res_ld %% 8");

  dac_conv_plan* pl = new (std::nothrow) dac_conv_plan();
  if (!pl) return set_error(-3, "out of host memory");
  ConvKParams& k = pl->kp;
  memset(&k, 0, sizeof(k));
  k.B = d->B; k.OH = d->OH; k.OW = d->OW; k.stride = d->stride;
  k.tile_h = d->tile_h; k.tile_w = d->tile_w;
  k.tiles_x = (d->OW + d->tile_w - 1) / d->tile_w;
  k.tiles_y = (d->OH + d->tile_h - 1) / d->tile_h;
  k.m_tiles = d->B * k.tiles_x * k.tiles_y;
  k.block_n = d->block_n;
  k.n_tiles = d->cout_pad / d->block_n;
  k.ngroups = d->ngroups; k.ntaps = d->ntaps;
  k.chunks0 = d->c0 / kChunkK; k.chunks1 = d->c1 / kChunkK; k.c0 = d->c0;
  k.per_image_w = d->per_image_w;
  k.b_bytes = (uint32_t)d->block_n * kChunkK * 2;
  memcpy(k.tap_dy, d->tap_dy, sizeof(k.tap_dy));
  memcpy(k.tap_dx, d->tap_dx, sizeof(k.tap_dx));
  k.epi = d->epi; k.act = d->act; k.cout = d->cout;
  k.bias = d->bias; k.bias_img = d->bias_img;
  k.film = d->film; k.film_ld = d->film_ld; k.film_off = d->film_off;
  k.ln_g = d->ln_g; k.ln_eps = d->ln_eps;
  k.res = static_cast<const __nv_bfloat16*>(d->res); k.res_ld = d->res_ld;
  k.res2 = static_cast<const __nv_bfloat16*>(d->res2); k.res2_ld = d->res2_ld;
  k.out = static_cast<__nv_bfloat16*>(d->out); k.out_ld = d->out_ld; k.out_coff = d->out_coff;
  k.res_f32 = d->res_f32; k.res_f32_ld = d->res_f32_ld;
  k.out_f32 = d->out_f32; k.out_f32_ld = d->out_f32_ld;
  k.out_scale = d->out_scale > 0 ? d->out_scale : 1;
  k.OHf = d->OH * k.out_scale; k.OWf = d->OW * k.out_scale;
  memcpy(k.out_oy, d->out_oy, 4); memcpy(k.out_ox, d->out_ox, 4);
  k.out_nchw = d->out_nchw; k.nchw_c = d->out_nchw_c; k.nchw_h = d->out_nchw_h; k.nchw_w = d->out_nchw_w;

  const uint32_t stage_bytes = kABytes + k.b_bytes;
  const int smem_budget = 227 * 1024 - 1024 /*align*/ - 256 /*barriers*/;
  int stages = smem_budget / (int)stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) { delete pl; return set_error(-2, "dac_conv_create: tile does not fit shared memory"); }
  k.stages = stages;
  pl->smem = stages * (int)stage_bytes + 1024 + 256;
  pl->tiles = k.ngroups * k.m_tiles * k.n_tiles;

  int rc = encode_act_map(&pl->mapA0, d->src0, d->c0, d->ld0, d->W, d->H, d->B, d->tile_w, d->tile_h, d->stride);
  if (rc == 0) {
    if (d->c1 > 0)
      rc = encode_act_map(&pl->mapA1, d->src1, d->c1, d->ld1, d->W, d->H, d->B, d->tile_w, d->tile_h, d->stride);
    else
      pl->mapA1 = pl->mapA0;
  }
  if (rc == 0) {
    PFN_encodeTiled enc = get_encode_fn();
    const int ctot = d->c0 + d->c1;
    const long long Z = (long long)d->ngroups * d->ntaps * (d->per_image_w ? d->B : 1);
    cuuint64_t dims[3] = {(cuuint64_t)ctot, (cuuint64_t)d->cout_pad, (cuuint64_t)Z};
    cuuint64_t strides[2] = {(cuuint64_t)ctot * 2, (cuuint64_t)d->cout_pad * ctot * 2};
    cuuint32_t box[3] = {(cuuint32_t)kChunkK, (cuuint32_t)d->block_n, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&pl->mapW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(d->weight), dims, strides,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) rc = set_error(-11, "cuTensorMapEncodeTiled(weight) failed: CUresult %d", (int)r);
  }
  if (rc != 0) { delete pl; return rc; }

  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  pl->grid = pl->tiles < sms ? pl->tiles : sms;
  static std::once_flag attr_once;
  std::call_once(attr_once, [] {
    cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  *out = pl;
  return 0;
}

extern "C" int dac_conv_launch(dac_conv_t pl, dac_stream_t stream) {
  if (!pl) return set_error(-1, "dac_conv_launch: null plan");
  conv_igemm_kernel<<<pl->grid, kThreads, pl->smem, static_cast<cudaStream_t>(stream)>>>(pl->mapA0, pl->mapA1,
                                                                                         pl->mapW, pl->kp);
  return check_launch("conv_igemm_kernel");
}

extern "C" void dac_conv_destroy(dac_conv_t pl) { delete pl; }

extern "C" int dac_conv_info(dac_conv_t pl, int32_t* tiles, int32_t* ctas, int32_t* smem_bytes, int32_t* stages) {
  if (!pl) return set_error(-1, "dac_conv_info: null plan");
  if (tiles) *tiles = pl->tiles;
  if (ctas) *ctas = pl->grid;
  if (smem_bytes) *smem_bytes = pl->smem;
  if (stages) *stages = pl->kp.stages;
  return 0;
}
