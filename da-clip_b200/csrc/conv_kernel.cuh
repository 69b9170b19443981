// Device side of the implicit-GEMM convolution / linear layer on tcgen05 tensor cores (sm_100a).
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
//   warps 2-9   epilogue, two groups of four warps, group g draining TMEM accumulator stage g (every 2nd tile):
//               tcgen05.ld the accumulator (one pixel per thread), apply bias / FiLM / SiLU|GELU / GEGLU /
//               channel-LayerNorm / q-softmax / residuals in fp32, store NHWC bf16 (and/or fp32, or fp32 NCHW for
//               final_conv).  Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
// The epilogue flavour is a template parameter (separate small kernels: the instruction cache matters - a single
// kernel with every flavour behind runtime flags was 150 KB of SASS and stalled 16 % of the time on fetch).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/dac_b200.h"
#include "ptx.cuh"
#include "tile_common.cuh"

namespace dac {

enum { KE_PLAIN = 0, KE_GEGLU = 1, KE_LN = 2, KE_QKV = 3, KE_NCHW = 4, KE_KVCTX = 5, KE_F32 = 6 };
constexpr int kKvPitch = 40;                                   // bf16 per row of a [128 px][32 ch] head tile
constexpr uint32_t kKvTileBytes = kTileM * kKvPitch * 2;         // 10 KB
constexpr uint32_t kKvStageBytes = 4 * kKvTileBytes;             // P and V head tiles, double-buffered: per group
constexpr int kCtxRecord = 32 * 32 + 64;                        // {C[32][32], m[32], S[32]} per (image, head)

// Division by a launch-time constant without the ~60-cycle IDIV sequence (Granlund-Montgomery round-up method):
// t = umulhi(mul, n); q = (t + ((n - t) >> s1)) >> s2.  Exact for 0 <= n < 2^31.
struct FastDiv {
  uint32_t mul, s1, s2, d;
};
__device__ __forceinline__ int fast_div(int n, const FastDiv& f) {
  const uint32_t t = __umulhi(f.mul, static_cast<uint32_t>(n));
  return static_cast<int>((t + ((static_cast<uint32_t>(n) - t) >> f.s1)) >> f.s2);
}
inline FastDiv make_fast_div(uint32_t d) {
  FastDiv f;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  f.mul = static_cast<uint32_t>(((1ull << 32) * ((1ull << l) - d)) / d + 1);
  f.s1 = l < 1 ? l : 1;
  f.s2 = l > 0 ? l - 1 : 0;
  f.d = d;
  return f;
}

struct ConvKParams {
  int B, OH, OW, stride;
  int tile_h, tile_w, tiles_x, tiles_y, m_tiles;
  int tile_w_shift;
  FastDiv fd_ntiles, fd_mtiles, fd_tx, fd_ty;
  FastDiv fd_mpairs;       // CTA-pair mode: m_tiles / 2 (pairs of M tiles per parity group)
  int n_tiles, block_n, ngroups, ntaps;
  int chunks0, chunks1, c0;
  int per_image_w;
  int stages;
  uint32_t b_bytes;        // one weight tile: block_n rows x 128 B
  uint32_t a_bytes;        // one activation load: (tile_h + ndy - 1) * tile_w rows x 128 B  (halo: (th+2)(tw+2) rows)
  uint32_t a_slot;         // a_bytes rounded up to the 1024 B swizzle atom: where the streamed weight tiles start
  uint32_t a_sbo;          // byte stride between the 8-pixel row groups of the activation operand (1024; halo: 1280)
  uint16_t tap_off[16];    // descriptor start offset (16 B units) of tap i within the activation load
  int halo;                // 1 / 2: nine taps from one haloed load (2: descriptor base_offset from the address)
  uint32_t b_res_bytes;    // > 0: the whole weight tensor stays resident in shared memory
  int r_chunks0, r_chunks1;   // fused 1x1 skip conv: K chunks of its two sources (0, 0 = none)
  uint32_t r_a_bytes;         // its activation tile: tile_h * tile_w rows x 128 B
  uint32_t r_b_bytes;         // its weight tile: block_n rows x 128 B (pair mode: the 64 real output channels)
  int ndy, ncols;
  int8_t col_dx[4][16];
  int8_t col_dy0[4][16];
  int8_t col_tap[4][16];
  // epilogue
  int cout;
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
  uint32_t stg_bytes;      // > 0: bf16 output goes through a swizzled shared-memory tile and a TMA store
  __nv_bfloat16* out_planar;   // QKV: planar k|v output [B][256][OH][OW]
  float* stats_out;            // PLAIN: per-pixel {mean, rstd} of the stored bf16 row (folded PreNorm)
  float stats_eps;
  const float* ln_stats;       // QKV: per-pixel {mean, rstd} of the input row
  const float* ln_colsum;      // QKV: sum_c W'[n][c]
  int stg_count;           // 1 or 2 staging tiles (2: the store of tile i overlaps the epilogue of tile i+1)
  int res_tma;             // the bf16 residual tile is TMA-loaded into the staging tile and updated in place
  int film_tmem;           // FiLM (scale + 1 | shift) of the current image lives in TMEM columns [bn, 3 bn) of the stage
  int pair;                // pixel-pair mode (see the issuer): the tensors are [B, H, W/2, 2C] views, block_n = 2 cout
  int film_cols;           // pair: the FiLM vectors have block_n / 2 entries and serve both pixels of a pair
  const float* kv_shift;   // KVCTX: [128] upper bound of k per channel, times log2(e) (host side: reduced to kv_shift_max)
  float kv_shift_max[4];   // KVCTX: the largest bound of each head - one scalar shift per head is all the softmax needs
  float* ctx_acc;          // KVCTX: [B][4][ctx_slots][kCtxRecord] fp32 partial records
  int ctx_slots, ctx_tpi;  // slots per (image, head); tiles per image
  int bias_sh;             // the tile's bias columns are staged in shared memory before the accumulator is awaited (PLAIN / F32 / GEGLU)
  int cta2;                // CTA-pair mode (cta_group::2): see the kernel; block_n, b_bytes describe the WHOLE / HALF weight tile
  int dbg;                 // profiling only (DAC_EPI_DEBUG, tools/prof_conv.py): 1 skip stores, 2 skip activation, 4 skip FiLM
};

struct TileCoord {
  int g, nt, n, y0, x0;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvKParams& p, int tile) {
  TileCoord t;
  const int rest = fast_div(tile, p.fd_ntiles);
  t.nt = tile - rest * p.n_tiles;
  t.g = fast_div(rest, p.fd_mtiles);
  const int mt = rest - t.g * p.m_tiles;
  const int r2 = fast_div(mt, p.fd_tx);
  const int tx = mt - r2 * p.tiles_x;
  t.n = fast_div(r2, p.fd_ty);
  const int ty = r2 - t.n * p.tiles_y;
  t.y0 = ty * p.tile_h;
  t.x0 = tx * p.tile_w;
  return t;
}

// v += 32 floats of this tile's staged bias (broadcast reads: every thread of the warp reads the same 16 bytes)
__device__ __forceinline__ void chunk_add_sh(const float* sh, float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 a = *reinterpret_cast<const float4*>(sh + 4 * q);
    v[4 * q] += a.x; v[4 * q + 1] += a.y; v[4 * q + 2] += a.z; v[4 * q + 3] += a.w;
  }
}

template <int ACT>
__device__ __forceinline__ float apply_act(float v) {
  if (ACT == DAC_ACT_SILU) return silu_f(v);
  if (ACT == DAC_ACT_GELU) return gelu_f(v);
  return v;
}

// Epilogue of one 128 x block_n accumulator tile.  Thread = one output pixel (TMEM lane `row`); the warp pair of
// the four warps of an epilogue group cover the four TMEM lane quadrants; a thread handles every column of its row.
template <int EPI, int ACT, bool FILM>
__device__ __forceinline__ void epilogue_tile(const ConvKParams& p, const TileCoord& t, uint32_t tmem_acc, int row,
                                              const float* film_sh, uint8_t* stg) {
  const int ty = row >> p.tile_w_shift, tx = row & (p.tile_w - 1);
  const int y = t.y0 + ty, x = t.x0 + tx;
  const bool valid = (y < p.OH) && (x < p.OW);
  const int Y = y * p.out_scale + p.out_oy[t.g], X = x * p.out_scale + p.out_ox[t.g];
  const long long opix = (static_cast<long long>(t.n) * p.OHf + Y) * p.OWf + X;
  const int n = t.n;
  float v[32];

  if (EPI == KE_NCHW) {
    // final_conv: 16-column tile, fp32 planar output cropped to the un-padded image; one warp per quadrant works.
    // (pixel-pair mode: two 16-column halves, the even and the odd pixel of the pair)
    for (int hp = 0; hp <= p.pair; ++hp) {
      uint32_t r[16];
      tmem_ld16(tmem_acc + 16 * hp, r);
      tmem_ld_wait();
      const int Xp = p.pair ? 2 * X + hp : X;
      if (valid && Y < p.nchw_h && Xp < p.nchw_w) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (j < p.nchw_c) {
            const float o = __uint_as_float(r[j]) + (p.bias ? __ldg(p.bias + j) : 0.f);
            p.out_nchw[((static_cast<long long>(n) * p.nchw_c + j) * p.nchw_h + Y) * p.nchw_w + Xp] = o;
          }
        }
      }
    }
    return;
  }

  if (EPI == KE_LN) {
    const int C = p.cout;
    if (C == 64) {
      // the whole row (2 chunks) lives in registers: ONE sweep of TMEM loads (issued back to back, one wait),
      // mean and centred variance from registers, then normalise / residual / store
      float w[2][32];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        {
          uint32_t r[32];
          tmem_ld32(tmem_acc + k * 32, r);
#pragma unroll
          for (int j = 0; j < 32; ++j) w[k][j] = __uint_as_float(r[j]);
        }
      }
      tmem_ld_wait();
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        {
          if (p.bias) chunk_add_f32(p.bias + k * 32, w[k]);
#pragma unroll
          for (int j = 0; j < 32; ++j) sum += w[k][j];
        }
      }
      const float mean = sum / C;
      float ss = 0.f;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            w[k][j] -= mean;
            ss = fmaf(w[k][j], w[k][j], ss);
          }
        }
      }
      const float rstd = rsqrtf(ss / C + p.ln_eps);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        {
          const int c = k * 32;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(p.ln_g + c) + q);
            w[k][4 * q] *= rstd * g.x;
            w[k][4 * q + 1] *= rstd * g.y;
            w[k][4 * q + 2] *= rstd * g.z;
            w[k][4 * q + 3] *= rstd * g.w;
          }
          if (p.res_tma) chunk_add_staged(stg, row, c, w[k]);
          else if (valid && p.res) chunk_add_bf16(p.res + opix * p.res_ld + c, w[k]);
          if (stg) chunk_stage_bf16(stg, row, c, w[k]);
          else if (valid) chunk_store_bf16(p.out + opix * p.out_ld + p.out_coff + c, w[k]);
        }
      }
      return;
    }
    // wider rows (C = 128, 256): three sweeps over this row's TMEM columns
    float sum = 0.f;
    for (int c = 0; c < C; c += 32) {
      chunk_from_tmem(tmem_acc + c, v);
      if (p.bias) chunk_add_f32(p.bias + c, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) sum += v[j];
    }
    const float mean = sum / C;
    float ss = 0.f;
    for (int c = 0; c < C; c += 32) {
      chunk_from_tmem(tmem_acc + c, v);
      if (p.bias) chunk_add_f32(p.bias + c, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float d = v[j] - mean;
        ss = fmaf(d, d, ss);
      }
    }
    const float rstd = rsqrtf(ss / C + p.ln_eps);
    for (int c = 0; c < C; c += 32) {
      chunk_from_tmem(tmem_acc + c, v);
      if (p.bias) chunk_add_f32(p.bias + c, v);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(p.ln_g + c) + q);
        v[4 * q] = (v[4 * q] - mean) * rstd * g.x;
        v[4 * q + 1] = (v[4 * q + 1] - mean) * rstd * g.y;
        v[4 * q + 2] = (v[4 * q + 2] - mean) * rstd * g.z;
        v[4 * q + 3] = (v[4 * q + 3] - mean) * rstd * g.w;
      }
      if (p.res_tma) chunk_add_staged(stg, row, c, v);
      else if (valid && p.res) chunk_add_bf16(p.res + opix * p.res_ld + c, v);
      if (stg) chunk_stage_bf16(stg, row, c, v);
      else if (valid) chunk_store_bf16(p.out + opix * p.out_ld + p.out_coff + c, v);
    }
    return;
  }

  if (EPI == KE_F32) {
    // out_f32 = acc + bias + res_f32 (the ViT blocks' fp32 residual stream, updated in place), optionally a bf16 copy.
    // A thread's row is 3 KB away from its neighbour's: with 16-byte accesses every warp instruction touched 32 sectors
    // and used half of each (out_proj 54 us against 26 us for the same GEMM without the stream); 256-bit accesses move
    // whole 32-byte sectors.  (Prefetching the residual chunks two ahead in registers changed nothing: 56 us.)
    for (int c = 0; c < p.block_n; c += 32) {
      const int ch = t.nt * p.block_n + c;
      if (ch >= p.cout) break;  // warp-uniform (cout_pad > cout)
      chunk_from_tmem(tmem_acc + c, v);
      if (p.bias_sh) chunk_add_sh(film_sh + c, v);
      else if (p.bias) chunk_add_f32(p.bias + ch, v);
      if (valid) {
        const float* rs = p.res_f32 + opix * p.res_f32_ld + ch;
        float* os = p.out_f32 ? p.out_f32 + opix * p.out_f32_ld + ch : nullptr;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          U32x8 r = ldg256(rs + 8 * q);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[8 * q + j] += __uint_as_float(r.v[j]);
            r.v[j] = __float_as_uint(v[8 * q + j]);
          }
          if (os) stg256(os + 8 * q, r);
        }
        if (p.out && !stg && !(p.dbg & 1)) chunk_store_bf16(p.out + opix * p.out_ld + p.out_coff + ch, v);
      }
      if (stg && !(p.dbg & 1)) chunk_stage_bf16(stg, row, c, v);
    }
    return;
  }

  if (EPI == KE_GEGLU) {
    const int hn = p.block_n >> 1;
    for (int c = 0; c < hn; c += 32) {
      float g[32];
      chunk_from_tmem(tmem_acc + c, v);
      chunk_from_tmem(tmem_acc + hn + c, g);
      const int col = t.nt * p.block_n + c;  // column in the (permuted) weight / bias row order
      if (p.bias_sh) {
        chunk_add_sh(film_sh + c, v);
        chunk_add_sh(film_sh + hn + c, g);
      } else {
        chunk_add_f32(p.bias + col, v);
        chunk_add_f32(p.bias + col + hn, g);
      }
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        gelu2_f(g[j], g[j + 1]);
        unpack_f32x2(mul_f32x2(pack_f32x2(v[j], v[j + 1]), pack_f32x2(g[j], g[j + 1])), v[j], v[j + 1]);
      }
      if (stg) chunk_stage_bf16(stg, row, c, v);
      else if (valid) chunk_store_bf16(p.out + opix * p.out_ld + p.out_coff + t.nt * hn + c, v);
    }
    return;
  }

  // KE_PLAIN / KE_QKV
  float st_sum = 0.f, st_sq = 0.f;       // PLAIN + stats_out: moments of the stored (bf16-rounded) row
  float ln_mean = 0.f, ln_rstd = 1.f;    // QKV + ln_stats: folded PreNorm of the input row
  if (EPI == KE_QKV && p.ln_stats && valid) {
    const float2 ms = __ldg(reinterpret_cast<const float2*>(p.ln_stats) + opix);
    ln_mean = ms.x;
    ln_rstd = ms.y;
  }
  for (int c = 0; c < p.block_n; c += 32) {
    const int ch = t.nt * p.block_n + c;
    if (ch >= p.cout) break;  // warp-uniform (cout_pad > cout)
    chunk_from_tmem(tmem_acc + c, v);
    if (EPI == KE_QKV) {
      if (p.ln_stats) {   // W' x  ->  W' LN(x) = rstd * (W' x - mean * colsum(W'))
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 cs = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + ch) + q);
          v[4 * q] = (v[4 * q] - ln_mean * cs.x) * ln_rstd;
          v[4 * q + 1] = (v[4 * q + 1] - ln_mean * cs.y) * ln_rstd;
          v[4 * q + 2] = (v[4 * q + 2] - ln_mean * cs.z) * ln_rstd;
          v[4 * q + 3] = (v[4 * q + 3] - ln_mean * cs.w) * ln_rstd;
        }
      }
      if (t.nt == 0) {  // q: softmax over the 32 channels of one head, times dim_head^-0.5
        float m = v[0];
#pragma unroll
        for (int j = 1; j < 32; ++j) m = fmaxf(m, v[j]);
        float s = 0.f;
        const float ml = m * 1.4426950408889634f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = ex2_approx(fmaf(v[j], 1.4426950408889634f, -ml));
          s += v[j];
        }
        const float inv = __fdividef(0.17677669529663687f, s);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= inv;
      }
      if (t.nt == 0) {
        if (stg) chunk_stage_bf16(stg, row, c, v);
        else if (valid) chunk_store_bf16(p.out + opix * p.out_ld + p.out_coff + ch, v);
      } else if (stg) {
        // k / v: transposed staging [channel][pixel] (lanes = consecutive pixels: conflict-free 2-byte stores);
        // one TMA store writes the 128 pixel-contiguous channel rows into the planar tensor
        __nv_bfloat16* tp = reinterpret_cast<__nv_bfloat16*>(stg) + c * kTileM + row;
#pragma unroll
        for (int j = 0; j < 32; ++j) tp[j * kTileM] = __float2bfloat16(v[j]);
      } else if (valid) {
        __nv_bfloat16* dp = p.out_planar + ((static_cast<long long>(n) * 256 + (ch - 128)) * p.OH + y) * p.OW + x;
        const long long plane = static_cast<long long>(p.OH) * p.OW;
#pragma unroll
        for (int j = 0; j < 32; ++j) dp[j * plane] = __float2bfloat16(v[j]);
      }
      continue;
    }
    if (!FILM && p.bias_sh) chunk_add_sh(film_sh + c, v);
    else if (p.bias) chunk_add_f32(p.bias + ch, v);
    if (p.bias_img) chunk_add_f32(p.bias_img + static_cast<long long>(n) * p.cout + ch, v);
    if (FILM && (p.dbg & 4)) {
    } else if (FILM && p.film_tmem) {
      // (scale + 1, shift) of this image replicated in every TMEM lane: two 32-column loads issued back to back (one
      // wait), no shared-memory traffic (a broadcast LDS.128 is four wavefronts on the pipe that bounds these layers);
      // applied as packed fp32 pairs (FFMA2)
      const int fcols = p.film_cols ? p.film_cols : p.block_n;   // pair mode: both pixels of the pair share the vectors
      const int fc = p.film_cols ? (c & (p.film_cols - 1)) : c;
      uint32_t fa[32], fb[32];
      tmem_ld32(tmem_acc + p.block_n + fc, fa);
      tmem_ld32(tmem_acc + p.block_n + fcols + fc, fb);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 2)
        unpack_f32x2(fma_f32x2(pack_f32x2(v[j], v[j + 1]),
                               pack_f32x2(__uint_as_float(fa[j]), __uint_as_float(fa[j + 1])),
                               pack_f32x2(__uint_as_float(fb[j]), __uint_as_float(fb[j + 1]))),
                     v[j], v[j + 1]);
    } else if (FILM) {
      // (scale + 1, shift) of this image, staged in shared memory by the epilogue warps when the image changes
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 a = *reinterpret_cast<const float4*>(film_sh + c + 4 * q);
        const float4 b = *reinterpret_cast<const float4*>(film_sh + p.block_n + c + 4 * q);
        v[4 * q] = fmaf(v[4 * q], a.x, b.x);
        v[4 * q + 1] = fmaf(v[4 * q + 1], a.y, b.y);
        v[4 * q + 2] = fmaf(v[4 * q + 2], a.z, b.z);
        v[4 * q + 3] = fmaf(v[4 * q + 3], a.w, b.w);
      }
    }
    if (ACT == DAC_ACT_SILU && !(p.dbg & 2)) {
      // x * sigmoid(x) = h + h * tanh(h), h = x / 2: FMUL2, two MUFU.TANH, FFMA2 per pair of values
      const uint64_t half2 = pack_f32x2(0.5f, 0.5f);
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const uint64_t h2 = mul_f32x2(pack_f32x2(v[j], v[j + 1]), half2);
        float h0, h1;
        unpack_f32x2(h2, h0, h1);
        unpack_f32x2(fma_f32x2(h2, pack_f32x2(tanh_approx(h0), tanh_approx(h1)), h2), v[j], v[j + 1]);
      }
    } else if (ACT == DAC_ACT_GELU && !(p.dbg & 2)) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) gelu2_f(v[j], v[j + 1]);
    }
    if (p.r_chunks0) {   // fused res_conv: its product sits in the next block_n TMEM columns
      float r2[32];
      chunk_from_tmem(tmem_acc + p.block_n + c, r2);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += r2[j];
    }
    if (p.res_tma) chunk_add_staged(stg, row, c, v);   // residual tile already in shared memory (TMA)
    if (valid) {
      if (p.res_f32) chunk_add_f32(p.res_f32 + opix * p.res_f32_ld + ch, v);
      if (p.res && !p.res_tma) chunk_add_bf16(p.res + opix * p.res_ld + ch, v);
      if (p.res2) chunk_add_bf16(p.res2 + opix * p.res2_ld + ch, v);
      if (p.out_f32) chunk_store_f32(p.out_f32 + opix * p.out_f32_ld + ch, v);
      if (p.out && !stg && !(p.dbg & 1)) chunk_store_bf16(p.out + opix * p.out_ld + p.out_coff + ch, v);
    }
    if (stg && !(p.dbg & 1)) chunk_stage_bf16(stg, row, c, v);
    if (EPI == KE_PLAIN && p.stats_out) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float r = __bfloat162float(__float2bfloat16(v[j]));   // what the consumer will read
        st_sum += r;
        st_sq = fmaf(r, r, st_sq);
      }
    }
  }
  if (EPI == KE_PLAIN && p.stats_out && valid) {
    const float mean = st_sum / p.cout;
    const float var = fmaxf(st_sq / p.cout - mean * mean, 0.f);
    reinterpret_cast<float2*>(p.stats_out)[opix] = make_float2(mean, rsqrtf(var + p.stats_eps));
  }
}

// ---- KVCTX: the LinearAttention context reduced straight out of the k|v accumulator (module_util.py:170-177) ----
// Per epilogue warp: C[h][16 d rows][16 e cols] of every head plus the softmax denominators S[h][16 d rows],
// carried in registers across the tiles of one image and STORED as this (CTA, group)'s partial record when the image
// changes (slot = 2 * (CTA - first CTA of the image) + group: the fold kernel adds the slots in order, so the result does
// not depend on which CTA finishes first).
struct KvCtxAcc {
  float c[4][2][4];
  float s[4][4];
};
__device__ __forceinline__ void kvctx_zero(KvCtxAcc& a) {
#pragma unroll
  for (int h = 0; h < 4; ++h) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      a.c[h][0][j] = 0.f;
      a.c[h][1][j] = 0.f;
      a.s[h][j] = 0.f;
    }
  }
}
__device__ __forceinline__ void kvctx_flush(const ConvKParams& p, int img, int quad, int lane, int group,
                                            KvCtxAcc& a) {
  if (img < 0) return;
  const int d = 16 * (quad >> 1) + (lane >> 2), e0 = 16 * (quad & 1) + 2 * (lane & 3);
  const int total = p.B * p.ctx_tpi;
  const int slot = 2 * (static_cast<int>(blockIdx.x) - tile_owner(img * p.ctx_tpi, total, gridDim.x)) + group;
  float* base = p.ctx_acc + (static_cast<long long>(img) * 4 * p.ctx_slots + slot) * kCtxRecord;
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    float* hb = base + static_cast<long long>(h) * p.ctx_slots * kCtxRecord;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      *reinterpret_cast<float2*>(hb + d * 32 + e0 + 8 * i) = make_float2(a.c[h][i][0], a.c[h][i][1]);
      *reinterpret_cast<float2*>(hb + (d + 8) * 32 + e0 + 8 * i) = make_float2(a.c[h][i][2], a.c[h][i][3]);
    }
    if ((quad & 1) == 0 && (lane & 3) == 0) {   // every column of the ones-product holds S[d]
      hb[1056 + d] = a.s[h][0];
      hb[1056 + d + 8] = a.s[h][2];
    }
  }
  kvctx_zero(a);
}
// One 128-pixel tile: thread = pixel `row`.  Head by head, P = exp(k - shift) and v go to shared memory as
// [pixel][channel] bf16 tiles (pitch 40: conflict-free for the row-per-lane stores AND the ldmatrix reads); then warp
// (mt, nh) runs C[16 mt.., 16 nh..] += P^T V over the 128 pixels on mma.sync (A = P^T and B = V both via
// ldmatrix.trans) and S += P^T 1.  Two buffers per operand: one named barrier per head is enough.
__device__ __forceinline__ void kvctx_tile(const ConvKParams& p, const TileCoord& t, uint32_t tmem_acc, int row,
                                           int quad, int lane, int group, uint8_t* stg, KvCtxAcc& a,
                                           uint64_t* tmem_empty_bar) {
  const int ty = row >> p.tile_w_shift, tx = row & (p.tile_w - 1);
  const bool valid = (t.y0 + ty < p.OH) && (t.x0 + tx < p.OW);
  __nv_bfloat16* tiles = reinterpret_cast<__nv_bfloat16*>(stg);
  const int mt = quad >> 1, nh = quad & 1;
  float v[32];
  // folded PreNorm (raw input rows, gain-folded weights): W' LN(x) = rstd * (W' x - mean * colsum(W'))
  float ka = 1.4426950408889634f, kb = 0.f, va = 1.f, vb = 0.f;
  if (p.ln_stats && valid) {
    const float2 ms = __ldg(reinterpret_cast<const float2*>(p.ln_stats) +
                            (static_cast<long long>(t.n) * p.OH + t.y0 + ty) * p.OW + t.x0 + tx);
    va = ms.y;
    vb = -ms.x * ms.y;
    ka = va * 1.4426950408889634f;
    kb = vb * 1.4426950408889634f;
  }
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    __nv_bfloat16* Ph = tiles + (h & 1) * (kTileM * kKvPitch);
    __nv_bfloat16* Vh = tiles + (2 + (h & 1)) * (kTileM * kKvPitch);
    chunk_from_tmem(tmem_acc + h * 32, v);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float4 sh = make_float4(p.kv_shift_max[h], p.kv_shift_max[h], p.kv_shift_max[h], p.kv_shift_max[h]);
      if (p.ln_stats) {
        const float4 cs = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + h * 32) + q);
        sh.x = fmaf(cs.x, -kb, sh.x); sh.y = fmaf(cs.y, -kb, sh.y); sh.z = fmaf(cs.z, -kb, sh.z); sh.w = fmaf(cs.w, -kb, sh.w);
      }
      v[4 * q] = valid ? ex2_approx(fmaf(v[4 * q], ka, -sh.x)) : 0.f;
      v[4 * q + 1] = valid ? ex2_approx(fmaf(v[4 * q + 1], ka, -sh.y)) : 0.f;
      v[4 * q + 2] = valid ? ex2_approx(fmaf(v[4 * q + 2], ka, -sh.z)) : 0.f;
      v[4 * q + 3] = valid ? ex2_approx(fmaf(v[4 * q + 3], ka, -sh.w)) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<uint4*>(Ph + row * kKvPitch + q * 8) =
          make_uint4(pack_bf16(v[q * 8], v[q * 8 + 1]), pack_bf16(v[q * 8 + 2], v[q * 8 + 3]),
                     pack_bf16(v[q * 8 + 4], v[q * 8 + 5]), pack_bf16(v[q * 8 + 6], v[q * 8 + 7]));
    chunk_from_tmem(tmem_acc + 128 + h * 32, v);
    if (p.ln_stats) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 cs = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + 128 + h * 32) + q);
        v[4 * q] = fmaf(v[4 * q], va, cs.x * vb);
        v[4 * q + 1] = fmaf(v[4 * q + 1], va, cs.y * vb);
        v[4 * q + 2] = fmaf(v[4 * q + 2], va, cs.z * vb);
        v[4 * q + 3] = fmaf(v[4 * q + 3], va, cs.w * vb);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<uint4*>(Vh + row * kKvPitch + q * 8) =
          valid ? make_uint4(pack_bf16(v[q * 8], v[q * 8 + 1]), pack_bf16(v[q * 8 + 2], v[q * 8 + 3]),
                             pack_bf16(v[q * 8 + 4], v[q * 8 + 5]), pack_bf16(v[q * 8 + 6], v[q * 8 + 7]))
                : make_uint4(0, 0, 0, 0);
    if (h == 3) {   // accumulator fully read: hand the TMEM stage back before the last reduction
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty_bar);
    }
    asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      uint32_t af[4], vb[4];
      ldmatrix_x4_trans(af, Ph + (16 * ks + ((lane >> 4) & 1) * 8 + (lane & 7)) * kKvPitch + 16 * mt +
                                ((lane >> 3) & 1) * 8);
      ldmatrix_x4_trans(vb, Vh + (16 * ks + ((lane >> 3) & 1) * 8 + (lane & 7)) * kKvPitch + 16 * nh +
                                (lane >> 4) * 8);
      mma_bf16_16816(a.c[h][0], af, vb[0], vb[1]);
      mma_bf16_16816(a.c[h][1], af, vb[2], vb[3]);
      mma_bf16_16816(a.s[h], af, 0x3F803F80u, 0x3F803F80u);   // B = ones: row sums of P^T
    }
  }
}

// CTA2: the CTA-pair build of a flavour (a cubin that holds cta_group::2 instructions can only be launched in clusters of
// even size - "cluster misconfiguration" otherwise - so the 1-CTA kernels must not contain them)
template <int EPI, int ACT, bool FILM, bool CTA2 = false>
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                  const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapOut,
                  const __grid_constant__ CUtensorMap mapOut2, const __grid_constant__ CUtensorMap mapR0,
                  const __grid_constant__ CUtensorMap mapR1, const __grid_constant__ CUtensorMap mapWR,
                  const __grid_constant__ CUtensorMap mapRes, const __grid_constant__ ConvKParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024 B alignment is required by the 128B swizzle atoms (TMA write and UMMA read agree on address bits).
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // [resident weights][ring of stages: activation tile (+ ndy weight tiles when streamed)][barriers]
  const bool b_resident = p.b_res_bytes != 0;
  uint8_t* b_res = smem;
  uint8_t* ring = smem + p.b_res_bytes;
  const int r_chunks = p.r_chunks0 + p.r_chunks1;
  const uint32_t b_stream = b_resident ? 0u : static_cast<uint32_t>(p.ndy) * p.b_bytes;
  const uint32_t main_tx = p.a_bytes + b_stream;   // bytes landing per K step
  uint32_t stage_bytes = p.a_slot + b_stream;
  if (r_chunks && stage_bytes < p.r_a_bytes + p.r_b_bytes) stage_bytes = p.r_a_bytes + p.r_b_bytes;
  uint8_t* stg_base = p.stg_bytes ? ring + static_cast<size_t>(p.stages) * stage_bytes : nullptr;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + static_cast<size_t>(p.stages) * stage_bytes +
                                               static_cast<size_t>(p.stg_bytes) * p.stg_count);
  uint64_t* full = bars;
  uint64_t* empty = bars + kMaxStages;
  uint64_t* tmem_full = bars + 2 * kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* b_full = tmem_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 1);
  uint64_t* res_bar = b_full + 2;   // one per epilogue group: residual tile landed in the staging tile
  uint64_t* b_empty = b_full + 4;   // parity-group layers: every MMA on the resident weights of the previous group is done
  float* film_sh = reinterpret_cast<float*>(bars + 32);   // [2][block_n]: scale + 1 | shift of the current image
  const int chunks = p.chunks0 + p.chunks1;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.ngroups * p.m_tiles * p.n_tiles;
  // ---- CTA-pair mode (p.cta2; streamed-weight layers, which are bound by L2 -> SM ingest: every CTA of the 1-CTA kernel
  // pulls the whole weight tile of every K step).  The two CTAs of a 2-CTA cluster work on two M tiles of the same N tile
  // as ONE M = 256 tcgen05.mma.cta_group::2: each loads its own activation tile and HALF of the weight tile (rows
  // [block_n / 2 * rank, +block_n / 2)), i.e. a third less operand traffic per CTA at N = 256; the leader (rank 0) issues
  // every MMA and its commits arrive in both CTAs; each CTA drains its own 128 accumulator rows with the usual epilogue.
  // The loops below walk pair indices v = (pair of M tiles, N tile); tile_of(v) is this CTA's tile.
  constexpr bool cta2 = CTA2;      // (p.cta2 says the same; the template keeps the pair instructions out of the 1-CTA builds)
  const uint32_t crank = cta2 ? cluster_ctarank() : 0u;
  int tile_begin, tile_end;
  if (cta2) {
    const long long total_pairs = static_cast<long long>(p.ngroups) * (p.m_tiles >> 1) * p.n_tiles;
    const int cl = blockIdx.x >> 1, ncl = gridDim.x >> 1;
    tile_begin = static_cast<int>(total_pairs * cl / ncl);
    tile_end = static_cast<int>(total_pairs * (cl + 1) / ncl);
  } else {
    tile_range(total_tiles, tile_begin, tile_end);
  }
  auto tile_of = [&](int v) -> int {
    if (!cta2) return v;
    const int rest = fast_div(v, p.fd_ntiles);                    // (group, pair of M tiles), group-major like the tiles
    const int g = p.ngroups == 1 ? 0 : fast_div(rest, p.fd_mpairs);
    const int mp = rest - g * (p.m_tiles >> 1);
    return (g * p.m_tiles + 2 * mp + static_cast<int>(crank)) * p.n_tiles + (v - rest * p.n_tiles);
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapW);
    tma_prefetch_desc(&mapOut);
    tma_prefetch_desc(&mapOut2);
    if (p.r_chunks0) {
      tma_prefetch_desc(&mapR0);
      tma_prefetch_desc(&mapR1);
      tma_prefetch_desc(&mapWR);
    }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);                      // pair mode: only the leader's barrier is used (see the producer)
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], (kEpiWarps / 2) * (cta2 ? 2 : 1));   // ... and both CTAs' epilogue warps
    }
    mbar_init(b_full, 1);
    mbar_init(b_empty, 1);
    mbar_init(&res_bar[0], 1);
    mbar_init(&res_bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (cta2) {
      tmem_alloc_pair(tmem_slot, kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (cta2) cluster_sync_all();     // both CTAs' barriers are initialised before either signals the other
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: the next kernel of the stream may be scheduled from now on (its CTAs land on an SM
  // when ours exits and run their set-up); everything above this line and the resident weight load below touch nothing
  // an earlier kernel produces, so they overlap the tail of the previous kernel - the waits sit in the producer (before
  // the first activation load) and in the epilogue warps (before any parameter / residual read or output write).
  griddep_launch();
#ifdef DAC_DEBUG
  if (threadIdx.x == 0 && blockIdx.x == 0)
    printf("[conv] tiles=%d k_steps=%d stages=%d block_n=%d tmem_base=%08x OH=%d OW=%d out=%p cout=%d epi=%d\n",
           total_tiles, k_steps, p.stages, p.block_n, tmem_base, p.OH, p.OW, p.out, p.cout, EPI);
#endif

  if (warp == 0) {
    // ===================== TMA producer (one elected thread) =====================
    if (elect_one()) {
      // Resident weights: the whole weight tensor of ONE parity group (ntaps x chunks tiles) sits in shared memory, laid
      // out in the ORDER THE MMA LOOP CONSUMES IT (N tile, chunk, column group, tap within the group), so the issuer
      // walks it with one add per tap.  Tiles are numbered group-major, so a CTA's contiguous range changes group at
      // most ngroups - 1 times: the weights are reloaded then (after the issuer's last MMA on the old ones).
      int w_group = -1;
      uint32_t w_loads = 0;
      auto load_group_weights = [&](int g) {
        if (w_loads) mbar_wait(b_empty, (w_loads - 1) & 1);
        if constexpr (cta2) {
          // CTA-pair build: this CTA keeps its HALF of the rows of every weight tile; both CTAs' bytes complete on the leader's
          // barrier (the leader's issuer is the only reader of b_full)
          const uint32_t bfl = mapa_u32(smem_u32(b_full), 0);
          if (crank == 0) mbar_arrive_expect_tx(b_full, 2u * p.b_res_bytes);
          for (int nt = 0; nt < p.n_tiles; ++nt)
            for (int ck = 0; ck < chunks; ++ck)
              for (int jt = 0; jt < p.ntaps; ++jt)
                tma_load_3d_pair(b_res + static_cast<size_t>((nt * chunks + ck) * p.ntaps + jt) * p.b_bytes, &mapW, bfl,
                                 ck * kChunkK, nt * p.block_n + static_cast<int>(crank) * (p.block_n >> 1),
                                 g * p.ntaps + p.col_tap[g][jt]);
        } else {
        mbar_arrive_expect_tx(b_full, p.b_res_bytes);
        for (int nt = 0; nt < p.n_tiles; ++nt)
          for (int ck = 0; ck < chunks; ++ck)
            for (int jt = 0; jt < p.ntaps; ++jt)
              tma_load_3d(b_res + static_cast<size_t>((nt * chunks + ck) * p.ntaps + jt) * p.b_bytes, &mapW, b_full,
                          ck * kChunkK, nt * p.block_n, g * p.ntaps + p.col_tap[g][jt]);
        }
        w_group = g;
        ++w_loads;
      };
      if (b_resident) {
        if (p.pair) {
          // per (64-channel source slice, ky): ONE 192-row block [W(kx=2); W(kx=1); W(kx=0)] - the even and the odd
          // chunk of the slice read overlapping 128-row windows of it (see the issuer)
          if constexpr (cta2) {
            // CTA-pair mode: this CTA's own 192-row block of the rank-specific layout (blocks 3 + 3 rank + ky of the weight
            // tensor, see the issuer); both CTAs' bytes complete on the LEADER's barrier
            const uint32_t bfl = mapa_u32(smem_u32(b_full), 0);
            if (crank == 0) mbar_arrive_expect_tx(b_full, 2u * p.b_res_bytes);
            for (int s_ = 0; s_ < (chunks >> 1); ++s_)
              for (int ky = 0; ky < 3; ++ky)
                tma_load_3d_pair(b_res + static_cast<size_t>(s_ * 3 + ky) * (3u * (p.block_n >> 1) * 128u), &mapW, bfl,
                                 s_ * kChunkK, 0, 3 + 3 * static_cast<int>(crank) + ky);
          } else {
          mbar_arrive_expect_tx(b_full, p.b_res_bytes);
          for (int s_ = 0; s_ < (chunks >> 1); ++s_)
            for (int ky = 0; ky < 3; ++ky)
              tma_load_3d(b_res + static_cast<size_t>(s_ * 3 + ky) * (3u * (p.block_n >> 1) * 128u), &mapW, b_full,
                          s_ * kChunkK, 0, ky);
          }
          w_group = 0;
          ++w_loads;
        } else if (tile_begin < tile_end) {
          load_group_weights(decode_tile(p, tile_of(tile_begin)).g);
        }
      }
      int stage = 0;
      uint32_t phase = 0;
      griddep_wait();
      const uint32_t full_lead = cta2 ? mapa_u32(smem_u32(full), 0) : 0u;   // the leader's full[0] (shared::cluster address)
      const int half_n = p.block_n >> 1;
      const bool lean_steps = !b_resident && p.ncols <= 3 && p.ndy <= 3 && r_chunks == 0 && !p.pair;
      const bool tap_steps = p.ndy == 1 && p.ncols > 3 && p.ncols <= 16 && r_chunks == 0 && !p.pair && !p.halo;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const TileCoord t = decode_tile(p, tile_of(tile));
        if (b_resident && t.g != w_group) load_group_weights(t.g);
        const int xin = t.x0 * p.stride, yin = t.y0 * p.stride;
        const int zbase = (p.per_image_w ? t.n * p.ngroups * p.ntaps : 0) + t.g * p.ntaps;
        const int ncoord = t.nt * p.block_n;
        if (b_resident && p.ncols == 1) {
          // Resident weights and one (haloed) activation load per chunk - the pixel-pair and halo layers of levels 0-1: the
          // main K loop issues nothing but that load (the fused skip conv's loads follow below)
          const int xa = xin + p.col_dx[t.g][0], ya = yin + p.col_dy0[t.g][0];
          for (int ck = 0; ck < chunks; ++ck) {
            const bool first = ck < p.chunks0;
            const CUtensorMap* mapA = first ? &mapA0 : &mapA1;
            const int ccoord = (first ? ck : ck - p.chunks0) * kChunkK;
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = ring + static_cast<size_t>(stage) * stage_bytes;
            if (p.dbg & 8) {          // profiling only (DAC_EPI_DEBUG & 8): no activation loads - the MMA / barrier skeleton alone
              if (crank == 0) mbar_arrive(&full[stage]);
            } else if constexpr (cta2) {
              if (crank == 0) mbar_arrive_expect_tx(&full[stage], 2u * main_tx);
              tma_load_4d_pair(sa, mapA, full_lead + 8u * stage, ccoord, xa, ya, t.n);
            } else {
              mbar_arrive_expect_tx(&full[stage], main_tx);
              tma_load_4d(sa, mapA, &full[stage], ccoord, xa, ya, t.n);
            }
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        } else
        if (tap_steps) {
          // One tap per load and up to 16 loads per chunk (the 4x4 stride-2 Downsample convs: 16 strided loads of four
          // N <= 128 MMAs each, i.e. 192-256 tensor cycles per load - the ~850-cycle general loop bounded them): the three
          // tap tables of the tile's group travel in registers as packed bytes.
          uint32_t dxw[4], dyw[4], tpw[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            dxw[q] = dyw[q] = tpw[q] = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              dxw[q] |= static_cast<uint32_t>(static_cast<uint8_t>(p.col_dx[t.g][4 * q + e])) << (8 * e);
              dyw[q] |= static_cast<uint32_t>(static_cast<uint8_t>(p.col_dy0[t.g][4 * q + e])) << (8 * e);
              tpw[q] |= static_cast<uint32_t>(static_cast<uint8_t>(p.col_tap[t.g][4 * q + e])) << (8 * e);
            }
          }
          const int wrow = ncoord + (cta2 ? static_cast<int>(crank) * half_n : 0);
          for (int ck = 0; ck < chunks; ++ck) {
            const bool first = ck < p.chunks0;
            const CUtensorMap* mapA = first ? &mapA0 : &mapA1;
            const int ccoord = (first ? ck : ck - p.chunks0) * kChunkK;
#pragma unroll 4
            for (int j = 0; j < p.ncols; ++j) {
              const int sh = 8 * (j & 3);
              const uint32_t wx = j < 8 ? (j < 4 ? dxw[0] : dxw[1]) : (j < 12 ? dxw[2] : dxw[3]);
              const uint32_t wy = j < 8 ? (j < 4 ? dyw[0] : dyw[1]) : (j < 12 ? dyw[2] : dyw[3]);
              const uint32_t wt = j < 8 ? (j < 4 ? tpw[0] : tpw[1]) : (j < 12 ? tpw[2] : tpw[3]);
              const int xa = xin + static_cast<int8_t>(wx >> sh), ya = yin + static_cast<int8_t>(wy >> sh);
              const int zt = zbase + static_cast<int8_t>(wt >> sh);
              mbar_wait(&empty[stage], phase ^ 1);
              uint8_t* sa = ring + static_cast<size_t>(stage) * stage_bytes;
              if constexpr (cta2) {
                const uint32_t fb = full_lead + 8u * stage;
                if (crank == 0) mbar_arrive_expect_tx(&full[stage], 2u * main_tx);
                tma_load_4d_pair(sa, mapA, fb, ccoord, xa, ya, t.n);
                if (!b_resident) tma_load_3d_pair(sa + p.a_slot, &mapW, fb, ck * kChunkK, wrow, zt);
              } else {
                mbar_arrive_expect_tx(&full[stage], main_tx);
                tma_load_4d(sa, mapA, &full[stage], ccoord, xa, ya, t.n);
                if (!b_resident) tma_load_3d(sa + p.a_slot, &mapW, &full[stage], ck * kChunkK, wrow, zt);
              }
              if (++stage == p.stages) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
          continue;
        }
        if (lean_steps) {
          // Streamed-weight K loops with at most three loads per chunk and three taps per load (1x1 layers and every Linear of
          // the transformers and the ViTs: one load, one tap; 3x3 layers: three column loads of three taps): the tap tables
          // are read ONCE per tile and the loops are fully unrolled.  The general loop below costs ~250 instructions per
          // load - ~850 cycles of this single thread against the 384-512 cycles the four MMAs of a GEMM step take: the
          // PRODUCER bounded those layers (tensor pipe 39 % busy, ncu source page of the ViT in_proj GEMM).
          int xa[3], ya[3], zt[3][3];
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            xa[j] = xin + p.col_dx[t.g][j];
            ya[j] = yin + p.col_dy0[t.g][j];
#pragma unroll
            for (int i = 0; i < 3; ++i) zt[j][i] = zbase + p.col_tap[t.g][j * p.ndy + i];
          }
          const int wrow = ncoord + (cta2 ? static_cast<int>(crank) * half_n : 0);
          for (int ck = 0; ck < chunks; ++ck) {
            const bool first = ck < p.chunks0;
            const CUtensorMap* mapA = first ? &mapA0 : &mapA1;
            const int ccoord = (first ? ck : ck - p.chunks0) * kChunkK;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              if (j < p.ncols) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* sa = ring + static_cast<size_t>(stage) * stage_bytes;
                if constexpr (cta2) {
                  const uint32_t fb = full_lead + 8u * stage;
                  if (crank == 0) mbar_arrive_expect_tx(&full[stage], 2u * main_tx);
                  tma_load_4d_pair(sa, mapA, fb, ccoord, xa[j], ya[j], t.n);
#pragma unroll
                  for (int i = 0; i < 3; ++i)
                    if (i < p.ndy)
                      tma_load_3d_pair(sa + p.a_slot + static_cast<size_t>(i) * p.b_bytes, &mapW, fb, ck * kChunkK, wrow, zt[j][i]);
                } else {
                  mbar_arrive_expect_tx(&full[stage], main_tx);
                  tma_load_4d(sa, mapA, &full[stage], ccoord, xa[j], ya[j], t.n);
#pragma unroll
                  for (int i = 0; i < 3; ++i)
                    if (i < p.ndy)
                      tma_load_3d(sa + p.a_slot + static_cast<size_t>(i) * p.b_bytes, &mapW, &full[stage], ck * kChunkK, wrow,
                                  zt[j][i]);
                }
                if (++stage == p.stages) {
                  stage = 0;
                  phase ^= 1;
                }
              }
            }
          }
          continue;
        }
        if (!(b_resident && p.ncols == 1))
        for (int ck = 0; ck < chunks; ++ck) {
          const CUtensorMap* mapA = ck < p.chunks0 ? &mapA0 : &mapA1;
          const int ccoord = (ck < p.chunks0 ? ck : ck - p.chunks0) * kChunkK;
          for (int j = 0; j < p.ncols; ++j) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = ring + static_cast<size_t>(stage) * stage_bytes;
            if constexpr (cta2) {
              // this CTA's activation tile and its half of the weight rows, both signalling the leader's barrier
              // (the leader announces the bytes of BOTH CTAs - they load equal amounts; the peer only issues its loads: its
              // bytes may land before the leader's arrival, which merely takes the transaction count through zero while the
              // arrival is still pending)
              const uint32_t fb = full_lead + 8u * stage;
              if (crank == 0) mbar_arrive_expect_tx(&full[stage], 2u * main_tx);
              tma_load_4d_pair(sa, mapA, fb, ccoord, xin + p.col_dx[t.g][j], yin + p.col_dy0[t.g][j], t.n);
              if (!b_resident)
                for (int i = 0; i < p.ndy; ++i)
                  tma_load_3d_pair(sa + p.a_slot + static_cast<size_t>(i) * p.b_bytes, &mapW, fb, ck * kChunkK,
                                   ncoord + static_cast<int>(crank) * half_n, zbase + p.col_tap[t.g][j * p.ndy + i]);
              if (++stage == p.stages) {
                stage = 0;
                phase ^= 1;
              }
              continue;
            }
            mbar_arrive_expect_tx(&full[stage], main_tx);
            tma_load_4d(sa, mapA, &full[stage], ccoord, xin + p.col_dx[t.g][j], yin + p.col_dy0[t.g][j], t.n);
            if (!b_resident) {
              for (int i = 0; i < p.ndy; ++i)
                tma_load_3d(sa + p.a_slot + static_cast<size_t>(i) * p.b_bytes, &mapW, &full[stage], ck * kChunkK,
                            ncoord, zbase + p.col_tap[t.g][j * p.ndy + i]);
            }
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        // fused 1x1 skip conv: one un-shifted tile of each 64-channel chunk of its sources + the weight tile
        for (int rk = 0; rk < r_chunks; ++rk) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = ring + static_cast<size_t>(stage) * stage_bytes;
          if constexpr (cta2) {
            // CTA-pair build: this CTA's source tile and its half of the skip weight rows (r_b_bytes describes the half)
            const uint32_t fb = full_lead + 8u * stage;
            const int wr_rows = static_cast<int>(p.r_b_bytes >> 7);      // rows of 128 B
            if (crank == 0) mbar_arrive_expect_tx(&full[stage], 2u * (p.r_a_bytes + p.r_b_bytes));
            if (rk < p.r_chunks0) tma_load_4d_pair(sa, &mapR0, fb, rk * kChunkK, xin, yin, t.n);
            else tma_load_4d_pair(sa, &mapR1, fb, (rk - p.r_chunks0) * kChunkK, xin, yin, t.n);
            tma_load_3d_pair(sa + p.r_a_bytes, &mapWR, fb, (p.pair ? (rk >> 1) : rk) * kChunkK, static_cast<int>(crank) * wr_rows, 0);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          mbar_arrive_expect_tx(&full[stage], p.r_a_bytes + p.r_b_bytes);
          if (rk < p.r_chunks0) tma_load_4d(sa, &mapR0, &full[stage], rk * kChunkK, xin, yin, t.n);
          else tma_load_4d(sa, &mapR1, &full[stage], (rk - p.r_chunks0) * kChunkK, xin, yin, t.n);
          // pair mode: the even and the odd chunk of a 64-channel slice use the same weight columns
          tma_load_3d(sa + p.r_a_bytes, &mapWR, &full[stage], (p.pair ? (rk >> 1) : rk) * kChunkK, 0, 0);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer.  The issue loop is the critical path of the 64-channel layers (one
    // M128 x N64 x K16 MMA is only 32 tensor-pipe cycles).  ALL lanes run the loop and the descriptor arithmetic so
    // that every operand is provably warp-uniform and lives in uniform registers; only the tcgen05 instructions sit
    // under elect.sync (a whole-loop `if (elected)` made the compiler compute descriptors in vector registers and
    // pay two R2UR per operand, ~3x the MMA time). ==========
    if (!cta2 || crank == 0) {
      const uint32_t idesc = make_idesc_bf16(cta2 ? 2 * kTileM : kTileM, p.block_n);
      const uint64_t desc_fixed = make_sw128_desc(0);                 // every field except the start address
      const uint64_t desc_fixed_a = make_sw128_desc(0, p.a_sbo);
      const uint32_t ring_lo = (smem_u32(ring) & 0x3FFFF) >> 4;
      const uint32_t bres_lo = (smem_u32(b_res) & 0x3FFFF) >> 4;
      const uint32_t stage_lo = stage_bytes >> 4, a_lo = p.a_slot >> 4,
                     b_lo = p.b_bytes >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int w_group = -1;
      uint32_t w_loads = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        if (b_resident) {
          const int g = p.ngroups == 1 ? 0 : fast_div(fast_div(tile, p.fd_ntiles), cta2 ? p.fd_mpairs : p.fd_mtiles);
          if (g != w_group) {               // (next) parity group: its weights replace the resident ones
            if (w_group >= 0) {
              if (elect_one()) {
                if constexpr (cta2) umma_commit_pair(b_empty);   // both CTAs' producers reload their halves
                else umma_commit(b_empty);
              }
              __syncwarp();
            }
            mbar_wait(b_full, w_loads & 1);
            ++w_loads;
            w_group = g;
          }
        }
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        uint32_t accumulate = 0;
        const int nt = tile - fast_div(tile, p.fd_ntiles) * p.n_tiles;
        uint32_t b_run = bres_lo + nt * (chunks * p.ntaps) * b_lo;   // resident weights of this N tile, front to back
        if (p.pair) {
          // ---- pixel-pair mode.  A row of the tile is a PAIR of horizontally adjacent pixels: the activation tensors
          // are viewed as [B, H, W/2, 2C] (64-channel chunk 2s = the even pixels of source slice s, chunk 2s + 1 = the
          // odd ones) and the accumulator row holds both outputs, columns [0, cout) = pixel 2i, [cout, 2 cout) = 2i + 1.
          // An M128 x N64 MMA is bound by its 6 KB shared-memory operand fetch (48 cycles for 32 of math); here the
          // centre views feed N = 128 MMAs (8 KB in 64 cycles: balanced) against two overlapping windows of ONE
          // 192-row weight block [W(kx=2); W(kx=1); W(kx=0)]:
          //   even chunk, pair i     : rows [64, 192) = W(1) | W(0)  ->  x[2i] is the centre of pixel 2i, the left of 2i+1
          //   odd  chunk, pair i     : rows [0, 128)  = W(2) | W(1)  ->  x[2i+1] is the right of 2i, the centre of 2i+1
          //   odd  chunk, pair i - 1 : rows [128, 192) = W(0), N = 64 into columns [0, cout)      (left of pixel 2i)
          //   even chunk, pair i + 1 : rows [0, 64)    = W(2), N = 64 into columns [cout, 2 cout)  (right of pixel 2i+1)
          // = 224 instead of 288 fetch-bound cycles per 256 output pixels and K step.
          // CTA-pair build (cta_group::2, M = 256): CTA r supplies the B rows of output columns [N/2 r, +N/2), so each CTA
          // holds its own 192-row block per (slice, ky), [E; O; S0; S2] =
          //   rank 0: [W(1); W(2); W(0)[0:32]; W(2)[0:32]]     rank 1: [W(0); W(1); W(0)[32:64]; W(2)[32:64]]
          // E / O = this CTA's half of the even / odd centre window, S0 / S2 = its half of the W(0) / W(2) side windows
          // (ops.pack_conv_pair).  Per CTA and MMA the operand fetch drops from 4 + 4 KB to 4 + 2 KB (N = 128) and from
          // 4 + 2 to 4 + 1 KB (N = 64): the centre MMAs become math-bound, and the shared-memory pipe has room for the TMA
          // writes and the epilogue's staging traffic.
          const uint32_t idesc_half = make_idesc_bf16(cta2 ? 2 * kTileM : kTileM, p.block_n >> 1);
          const uint32_t half_cols = p.block_n >> 1;
          // (half_cols rows per kx block: 64 for the ResBlock layers, 16 for final_conv; every window starts on a multiple
          // of 8 rows = one 1024-byte swizzle atom)
          const uint32_t blk_lo = (3u * half_cols * 128u) >> 4, rows64_lo = (half_cols * 128u) >> 4;
          for (int ck = 0; ck < chunks; ++ck) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t a0 = ring_lo + stage * stage_lo;
            const uint32_t odd = ck & 1;
            const uint32_t wsrc = bres_lo + (ck >> 1) * 3 * blk_lo;
            for (int ky = 0; ky < 3; ++ky) {
              const uint32_t wb = wsrc + ky * blk_lo;
              // centre view (pair i): tap (ky, 1); side view: tap (ky, 0) for the odd chunk, (ky, 2) for the even one
              const uint32_t a_c = a0 + p.tap_off[ky * 3 + 1];
              const uint32_t a_s = a0 + p.tap_off[ky * 3 + (odd ? 0 : 2)];
              uint64_t adesc_c = desc_fixed_a | a_c, adesc_s = desc_fixed_a | a_s;
              if (p.halo == 2) {
                adesc_c |= static_cast<uint64_t>((a_c >> 3) & 7) << 49;
                adesc_s |= static_cast<uint64_t>((a_s >> 3) & 7) << 49;
              }
              const uint64_t bdesc_c = desc_fixed | (cta2 ? (odd ? wb + rows64_lo : wb) : (odd ? wb : wb + rows64_lo));
              const uint64_t bdesc_s = desc_fixed | (cta2 ? (odd ? wb + 2 * rows64_lo : wb + 2 * rows64_lo + (rows64_lo >> 1))
                                                          : (odd ? wb + 2 * rows64_lo : wb));
              const uint32_t d_s = d_tmem + (odd ? 0u : half_cols);
              if (elect_one()) {
                if constexpr (cta2) {
                  umma_bf16_pair(d_tmem, adesc_c, bdesc_c, idesc, accumulate);
                  umma_bf16_pair(d_tmem, adesc_c + 2, bdesc_c + 2, idesc, 1u);
                  umma_bf16_pair(d_tmem, adesc_c + 4, bdesc_c + 4, idesc, 1u);
                  umma_bf16_pair(d_tmem, adesc_c + 6, bdesc_c + 6, idesc, 1u);
                  umma_bf16_pair(d_s, adesc_s, bdesc_s, idesc_half, 1u);
                  umma_bf16_pair(d_s, adesc_s + 2, bdesc_s + 2, idesc_half, 1u);
                  umma_bf16_pair(d_s, adesc_s + 4, bdesc_s + 4, idesc_half, 1u);
                  umma_bf16_pair(d_s, adesc_s + 6, bdesc_s + 6, idesc_half, 1u);
                } else {
                  umma_bf16(d_tmem, adesc_c, bdesc_c, idesc, accumulate);
                  umma_bf16(d_tmem, adesc_c + 2, bdesc_c + 2, idesc, 1u);
                  umma_bf16(d_tmem, adesc_c + 4, bdesc_c + 4, idesc, 1u);
                  umma_bf16(d_tmem, adesc_c + 6, bdesc_c + 6, idesc, 1u);
                  umma_bf16(d_s, adesc_s, bdesc_s, idesc_half, 1u);
                  umma_bf16(d_s, adesc_s + 2, bdesc_s + 2, idesc_half, 1u);
                  umma_bf16(d_s, adesc_s + 4, bdesc_s + 4, idesc_half, 1u);
                  umma_bf16(d_s, adesc_s + 6, bdesc_s + 6, idesc_half, 1u);
                }
              }
              accumulate = 1u;
            }
            if (elect_one()) {
              if constexpr (cta2) umma_commit_pair(&empty[stage]);
              else umma_commit(&empty[stage]);
            }
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        } else if (p.ndy == 1 && !p.halo) {
          // One (activation, weight) tile pair per step - GEMM-shaped K loops and the one-tap-per-load convs; see the producer
          const uint32_t t_off = p.tap_off[0];
          const int steps = chunks * p.ncols;
          for (int ck = 0; ck < steps; ++ck) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t a0 = ring_lo + stage * stage_lo;
            const uint64_t adesc = desc_fixed_a | (a0 + t_off);
            const uint64_t bdesc = desc_fixed | (b_resident ? b_run : a0 + a_lo);
            b_run += b_lo;
            if (elect_one()) {
              if constexpr (cta2) {
                umma_bf16_pair(d_tmem, adesc, bdesc, idesc, accumulate);
                umma_bf16_pair(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
                umma_bf16_pair(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
                umma_bf16_pair(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
                umma_commit_pair(&empty[stage]);
              } else {
                umma_bf16(d_tmem, adesc, bdesc, idesc, accumulate);
                umma_bf16(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
                umma_bf16(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
                umma_bf16(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
                umma_commit(&empty[stage]);
              }
            }
            accumulate = 1u;
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        } else
        for (int ck = 0; ck < chunks; ++ck) {
          for (int j = 0; j < p.ncols; ++j) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t a0 = ring_lo + stage * stage_lo;
            uint32_t bs_run = a0 + a_lo;
            for (int i = 0; i < p.ndy; ++i) {
              // tap i of the load: the same activation tile, shifted down by whole tile rows (halo: and sideways
              // by one pixel = 128 B, with the row groups (tile_w + 2) * 128 B apart)
              const uint32_t a_tap = a0 + p.tap_off[i];
              uint64_t adesc = desc_fixed_a | a_tap;
              if (p.halo == 2) adesc |= static_cast<uint64_t>((a_tap >> 3) & 7) << 49;   // matrix base offset
              const uint64_t bdesc = desc_fixed | (b_resident ? b_run : bs_run);
              bs_run += b_lo;
              b_run += b_lo;
              if (elect_one()) {
                if constexpr (cta2) {
                  umma_bf16_pair(d_tmem, adesc, bdesc, idesc, accumulate);
                  umma_bf16_pair(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
                  umma_bf16_pair(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
                  umma_bf16_pair(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
                } else {
                  umma_bf16(d_tmem, adesc, bdesc, idesc, accumulate);
                  umma_bf16(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
                  umma_bf16(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
                  umma_bf16(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
                }
              }
              accumulate = 1u;
            }
            if (elect_one()) {
              if constexpr (cta2) umma_commit_pair(&empty[stage]);    // the stage is free again in BOTH CTAs
              else umma_commit(&empty[stage]);
            }
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        for (int rk = 0; rk < r_chunks; ++rk) {   // W_r . rsrc into TMEM columns [block_n, 2 block_n)
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a0 = ring_lo + stage * stage_lo;
          const uint64_t adesc = desc_fixed | a0;
          const uint64_t bdesc = desc_fixed | (a0 + (p.r_a_bytes >> 4));
          // pair mode: the 1x1 conv of the even pixels lands in the first half of the second accumulator, odd: second half
          const uint32_t d_r = d_tmem + p.block_n + ((p.pair && (rk & 1)) ? (p.block_n >> 1) : 0);
          const uint32_t idesc_r = p.pair ? make_idesc_bf16(cta2 ? 2 * kTileM : kTileM, p.block_n >> 1) : idesc;
          const uint32_t acc_r = p.pair ? (rk >= 2 ? 1u : 0u) : (rk ? 1u : 0u);
          if (elect_one()) {
            if constexpr (cta2) {
              umma_bf16_pair(d_r, adesc, bdesc, idesc_r, acc_r);
              umma_bf16_pair(d_r, adesc + 2, bdesc + 2, idesc_r, 1u);
              umma_bf16_pair(d_r, adesc + 4, bdesc + 4, idesc_r, 1u);
              umma_bf16_pair(d_r, adesc + 6, bdesc + 6, idesc_r, 1u);
              umma_commit_pair(&empty[stage]);
            } else {
              umma_bf16(d_r, adesc, bdesc, idesc_r, acc_r);
              umma_bf16(d_r, adesc + 2, bdesc + 2, idesc_r, 1u);
              umma_bf16(d_r, adesc + 4, bdesc + 4, idesc_r, 1u);
              umma_bf16(d_r, adesc + 6, bdesc + 6, idesc_r, 1u);
              umma_commit(&empty[stage]);
            }
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) {
          if constexpr (cta2) umma_commit_pair(&tmem_full[acc]);      // both CTAs' epilogue groups
          else umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue: two groups of four warps (2..5 / 6..9) =====================
    // Group g owns TMEM accumulator stage g, i.e. every second tile of this CTA: two tiles are in the epilogue at
    // once, so the latency chain of one (barrier wait, TMEM loads, residual loads, staging, TMA store) hides
    // behind the other - what bounds the short-K layers (1x1 convs, to_qkv, to_out).
    const int quad = warp & 3;          // TMEM lane quadrant this warp may access
    const int group = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int gthread = threadIdx.x - 64 - group * 128;   // 0..127 within the group
    float* film_g = film_sh + group * 512;
    uint8_t* stg = stg_base ? stg_base + group * p.stg_bytes : nullptr;
    uint32_t acc_phase = 0;
    int film_key = -1;
    griddep_wait();
    const uint32_t tmem_empty_lead = cta2 ? mapa_u32(smem_u32(&tmem_empty[group]), 0) : 0u;
    if (EPI == KE_KVCTX) {
      KvCtxAcc cacc;
      kvctx_zero(cacc);
      int cur_img = -1;
      for (int tile = tile_begin + group; tile < tile_end; tile += 2) {
        const TileCoord t = decode_tile(p, tile);
        if (t.n != cur_img) {
          kvctx_flush(p, cur_img, quad, lane, group, cacc);
          cur_img = t.n;
        }
        mbar_wait(&tmem_full[group], acc_phase);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + group * kAccStride;
        kvctx_tile(p, t, tmem_acc, row, quad, lane, group, stg, cacc, &tmem_empty[group]);
        acc_phase ^= 1;
      }
      kvctx_flush(p, cur_img, quad, lane, group, cacc);
    } else
    for (int tile = tile_begin + group; tile < tile_end; tile += 2) {
      const TileCoord t = decode_tile(p, tile_of(tile));
      if (FILM) {
        // FiLM parameters depend on (image, N tile) only: restage when that pair changes (rare with contiguous
        // tile ranges).  Named barrier 1 + group = the 128 threads of this group.
        const int key = t.n * p.n_tiles + t.nt;
        if (key != film_key && p.film_tmem) {
          // every thread writes the image's 2 x block_n parameters into its own TMEM lane, beside the accumulator
          film_key = key;
          const float* src = p.film + static_cast<long long>(t.n) * p.film_ld + p.film_off + t.nt * p.block_n;
          const uint32_t faddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + group * kAccStride + p.block_n;
          const int fcols = p.film_cols ? p.film_cols : p.block_n, fcout = p.film_cols ? p.film_cols : p.cout;
          for (int c = 0; c < 2 * fcols; c += 32) {
            const bool is_scale = c < fcols;
            const float* s_ = is_scale ? src + c : src + fcout + (c - fcols);
            uint32_t r[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(s_) + q);
              r[4 * q] = __float_as_uint(is_scale ? a.x + 1.0f : a.x);
              r[4 * q + 1] = __float_as_uint(is_scale ? a.y + 1.0f : a.y);
              r[4 * q + 2] = __float_as_uint(is_scale ? a.z + 1.0f : a.z);
              r[4 * q + 3] = __float_as_uint(is_scale ? a.w + 1.0f : a.w);
            }
            tmem_st32(faddr + c, r);
          }
          tmem_st_wait();
        } else if (key != film_key) {
          film_key = key;
          asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
          const float* src = p.film + static_cast<long long>(t.n) * p.film_ld + p.film_off + t.nt * p.block_n;
          // (pair mode: the 64-entry vectors serve both pixels of the pair - replicated across the two column halves)
          const int fcout = p.film_cols ? p.film_cols : p.cout;
          for (int i = gthread; i < 2 * p.block_n; i += 128) {
            const int c = i < p.block_n ? i : i - p.block_n;
            const int cc = p.film_cols ? (c & (p.film_cols - 1)) : c;
            film_g[i] = i < p.block_n ? __ldg(src + cc) + 1.0f : __ldg(src + fcout + cc);
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
        }
      }
      if (!FILM && p.bias_sh && t.nt != film_key) {
        // the bias columns of this N tile, staged once per (group, N tile) BEFORE the accumulator is awaited: the per-chunk
        // bias loads sat on the epilogue's critical chain (accumulator ready -> TMEM load -> L2 round trip per 32 columns ->
        // first add: 11 % of all stall samples of the ViT GEMMs)
        film_key = t.nt;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
        const int c0 = t.nt * p.block_n;
        for (int i = gthread; i < p.block_n; i += 128) film_g[i] = (c0 + i < p.cout) ? __ldg(p.bias + c0 + i) : 0.f;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
      }
      if (stg && gthread == 0) {
        // this group's previous TMA store must have finished reading the staging tile before it is rewritten
        tma_store_wait_read();
        if (p.res_tma) {   // fetch the residual tile now: it lands while the MMAs of this tile are still running
          const int cols = p.block_n;
          mbar_arrive_expect_tx(&res_bar[group], static_cast<uint32_t>(cols) * kTileM * 2);
          for (int s_ = 0; s_ * 64 < cols; ++s_)
            tma_load_4d(stg + s_ * (kTileM * 128), &mapRes, &res_bar[group], t.nt * cols + s_ * 64, t.x0, t.y0, t.n);
          // ... and pull the residual of this group's NEXT tile into L2: its load can only be issued once the store of
          // this tile has left the staging buffer, and a DRAM round trip at that point was the longest link of the
          // per-group chain (store drained -> residual lands -> epilogue -> store) - 4.5 us per tile, whatever its size
          if (tile + 2 < tile_end) {
            const TileCoord tn = decode_tile(p, tile_of(tile + 2));
            for (int s_ = 0; s_ * 64 < cols; ++s_)
              tma_prefetch_l2_4d(&mapRes, tn.nt * cols + s_ * 64, tn.x0, tn.y0, tn.n);
          }
        }
      }
      mbar_wait(&tmem_full[group], acc_phase);
      tc_fence_after();
      if (stg) {
        asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
        if (p.res_tma) mbar_wait(&res_bar[group], acc_phase);
      }
      const uint32_t tmem_acc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + group * kAccStride;
      epilogue_tile<EPI, ACT, FILM>(p, t, tmem_acc, row, film_g, stg);
      if (stg) {
        fence_proxy_async();                       // generic-proxy smem writes -> visible to the TMA engine
        asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
        if (gthread == 0) {
          const int cols = (EPI == KE_GEGLU) ? (p.block_n >> 1) : p.block_n;   // output columns of this tile
          if (EPI == KE_QKV && t.nt > 0) {
            tma_store_4d(&mapOut2, stg, t.x0, t.y0, (t.nt - 1) * 128, t.n);    // planar [x, y, channel, image]
          } else {
            for (int s_ = 0; s_ * 64 < cols; ++s_)
              tma_store_4d(&mapOut, stg + s_ * (kTileM * 128), p.out_coff + t.nt * cols + s_ * 64, t.x0, t.y0, t.n);
          }
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (cta2) mbar_arrive_cluster(tmem_empty_lead);   // the leader issues for both CTAs: its barrier collects both
        else mbar_arrive(&tmem_empty[group]);
      }
      acc_phase ^= 1;
    }
    if (EPI != KE_KVCTX && stg && gthread == 0) tma_store_wait_read();   // smem must outlive the last bulk store's reads
  }

  tc_fence_before();
  if constexpr (cta2) cluster_sync_all();     // the peer's shared memory and barriers stay alive until both CTAs are done
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (cta2) tmem_dealloc_pair(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

typedef void (*ConvKernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                             const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                             const CUtensorMap, const ConvKParams);

// The epilogue flavours that exist as separate kernels; everything else in the epilogue is a warp-uniform
// runtime branch on a pointer.
inline ConvKernelFn pick_conv_kernel(int epi, int act, bool film, bool nchw, bool f32_stream = false, bool cta2 = false) {
  if (cta2) {   // CTA-pair builds of the flavours the streamed-weight layers use
    if (nchw || epi == DAC_EPI_QKV || epi == DAC_EPI_KVCTX) return nullptr;
    if (f32_stream && epi == DAC_EPI_PLAIN && act == DAC_ACT_NONE && !film) return conv_igemm_kernel<KE_F32, DAC_ACT_NONE, false, true>;
    if (epi == DAC_EPI_GEGLU) return conv_igemm_kernel<KE_GEGLU, DAC_ACT_NONE, false, true>;
    if (epi == DAC_EPI_LN) return conv_igemm_kernel<KE_LN, DAC_ACT_NONE, false, true>;
    if (act == DAC_ACT_SILU)
      return film ? conv_igemm_kernel<KE_PLAIN, DAC_ACT_SILU, true, true> : conv_igemm_kernel<KE_PLAIN, DAC_ACT_SILU, false, true>;
    if (act == DAC_ACT_GELU && !film) return conv_igemm_kernel<KE_PLAIN, DAC_ACT_GELU, false, true>;
    if (act == DAC_ACT_NONE && !film) return conv_igemm_kernel<KE_PLAIN, DAC_ACT_NONE, false, true>;
    return nullptr;
  }
  if (nchw) return conv_igemm_kernel<KE_NCHW, DAC_ACT_NONE, false>;
  if (f32_stream && epi == DAC_EPI_PLAIN && act == DAC_ACT_NONE && !film) return conv_igemm_kernel<KE_F32, DAC_ACT_NONE, false>;
  if (epi == DAC_EPI_GEGLU) return conv_igemm_kernel<KE_GEGLU, DAC_ACT_NONE, false>;
  if (epi == DAC_EPI_LN) return conv_igemm_kernel<KE_LN, DAC_ACT_NONE, false>;
  if (epi == DAC_EPI_QKV) return conv_igemm_kernel<KE_QKV, DAC_ACT_NONE, false>;
  if (epi == DAC_EPI_KVCTX) return conv_igemm_kernel<KE_KVCTX, DAC_ACT_NONE, false>;
  if (act == DAC_ACT_SILU)
    return film ? conv_igemm_kernel<KE_PLAIN, DAC_ACT_SILU, true> : conv_igemm_kernel<KE_PLAIN, DAC_ACT_SILU, false>;
  if (act == DAC_ACT_GELU && !film) return conv_igemm_kernel<KE_PLAIN, DAC_ACT_GELU, false>;
  if (act == DAC_ACT_NONE && !film) return conv_igemm_kernel<KE_PLAIN, DAC_ACT_NONE, false>;
  return nullptr;
}

}  // namespace dac
