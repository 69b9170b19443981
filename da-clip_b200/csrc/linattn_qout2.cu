// LinearAttention, query side for 64-channel inputs, second generation (round 2): ONE chained-GEMM kernel that takes
// the RAW activation tensor (module_util.py:89-97 PreNorm, :77-86 LayerNorm, :170-185 of the reference):
//
//   q    = softmax_channels-of-head(W_q LN(x)) * 32^-0.5            GEMM 1: [128 px, 64] x [64, 128] on the raw tile
//   out  = LayerNorm_c(W_eff[b] q + bias) * g + x                   GEMM 2: [128 px, 128] x [128, 64]; x = the same raw tile
//
// PreNorm is folded into GEMM 1: with row-centred weights W_qc = W_q - rowmean(W_q) (gain already folded in) the mean
// term vanishes, W_q LN(x) = rstd (W_qc x), and rstd - one scalar per accumulator row - rides inside the FFMA that feeds
// exp2.  Four warps compute the row moments of each landed tile (8 LDS.128 + ~70 packed fp32 ops per row); no normalised
// tile exists anywhere.  Differences from round 1's linattn_qout_kernel<64> (linattn_qout.cu):
//   * the `prenorm` pass (56 us at 256^2, batch 16) and the normalised tensor are gone;
//   * the residual is the raw tile itself, still sitting in its pipeline stage: the tensor is read ONCE per tile (round
//     1 read the normalised tensor through one map and the raw one through another: 269 MB per launch instead of 134);
//   * 16 epilogue warps instead of 8: each in-flight tile is shared by two 4-warp groups - group s softmaxes heads 2s,
//     2s+1 (64 of the 128 q columns) and finishes 32 of the 64 output channels; the row statistics of the output
//     LayerNorm meet in shared memory (one named barrier per tile);
//   * fewer instructions (the kernel is bound by instruction issue, ncu: 56 % of the issue slots with every pipe below
//     35 %): the softmax shift is the data-independent bound c_h = max_d ||W_qc[d]|| sqrt(C) >= |q_d| instead of the row
//     maximum (softmax is shift-invariant; exp stays in range while c_h <= 40), bias and LayerNorm gain come from the
//     kernel-parameter constant bank as instruction operands, the arithmetic is packed fp32x2 throughout.
// Roles (22 warps): warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 row statistics, warps 6-21 epilogue: tile pair
// t = i & 1, half s.  TMEM: [256 t, +128) GEMM-1 accumulator, [256 t + 128, +64) GEMM-2 accumulator.
#include <cuda.h>
#include <cuda_runtime.h>
#include <new>
#include <type_traits>

#include "../../include/dac_b200.h"
#include "common.h"
#include "linattn_qout_common.h"
#include "linattn_rowstats.cuh"
#include "tensormap.h"
#include "tile_common.cuh"

namespace dac {

constexpr uint32_t kQ2Slab = kTileM * 128;   // 128 rows x 64 bf16 (16 KB)
constexpr int kQ2Threads = 704;
constexpr int kQ2XStages = 6;                // raw activation tiles (held until the residual has been added)
constexpr int kQ2WStages = 2;                // W_eff[b]: both 64-column K chunks of one tile per stage (2 x 8 KB)

__global__ void __launch_bounds__(kQ2Threads, 1)
linattn_qout2_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapWq,
                     const __grid_constant__ CUtensorMap mapWeff, const __grid_constant__ CUtensorMap mapOut,
                     const __grid_constant__ Qout2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wq = smem;                                   // [128 rows x 128 B], resident
  uint8_t* xring = wq + kQ2Slab;                        // [kQ2XStages] raw tiles: A operand of GEMM 1 and the residual
  uint8_t* wring = xring + kQ2XStages * kQ2Slab;        // [kQ2WStages] x (2 x [64 rows x 128 B])
  uint8_t* a2 = wring + kQ2WStages * kQ2Slab;           // [2 tile pairs][2 slabs]: q tile; slab 0 doubles as output staging
  float* rowa = reinterpret_cast<float*>(a2 + 4 * kQ2Slab);               // [kQ2XStages][128]: rstd log2e of the pixel
  float* stat = rowa + kQ2XStages * kTileM;                               // [2 tile pairs][2 halves][128 rows][2]: {sum, sumsq}
  uint64_t* bars = reinterpret_cast<uint64_t*>(stat + 2 * 2 * kTileM * 2);
  uint64_t* x_full = bars;              // [8]
  uint64_t* x_empty = bars + 8;         // [8]  residual consumed (one arrival per tile)
  uint64_t* st_full = bars + 16;        // [8]  row statistics written (count 128)
  uint64_t* w_full = bars + 24;         // [2]
  uint64_t* w_empty = bars + 26;        // [2]  GEMM 2 has read the stage
  uint64_t* acc1_full = bars + 28;      // [2]
  uint64_t* acc1_empty = bars + 30;     // [2]  count 256
  uint64_t* a2_full = bars + 32;        // [2]  count 256
  uint64_t* d2_full = bars + 34;        // [2]
  uint64_t* wq_full = bars + 36;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 37);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int begin, end;
  tile_range(p.tiles, begin, end);
  const int n = end - begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapX);
    tma_prefetch_desc(&mapWq);
    tma_prefetch_desc(&mapWeff);
    tma_prefetch_desc(&mapOut);
    for (int s = 0; s < kQ2XStages; ++s) {
      mbar_init(&x_full[s], 1);
      mbar_init(&x_empty[s], 1);
      mbar_init(&st_full[s], 128);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
      mbar_init(&acc1_full[s], 1);
      mbar_init(&acc1_empty[s], 256);
      mbar_init(&a2_full[s], 256);
      mbar_init(&d2_full[s], 1);
    }
    mbar_init(wq_full, 1);
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
  griddep_launch();   // see conv_kernel.cuh: set-up and the W_q load overlap the previous kernel's tail

  if (warp == 0) {
    // ===================== TMA producer: per tile the raw activation tile and both K chunks of W_eff[image] =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(wq_full, kQ2Slab);
      tma_load_2d(wq, &mapWq, wq_full, 0, 0);
      griddep_wait();
      int xs = 0, ws = 0;
      uint32_t xph = 0, wph = 0;
      for (int i = 0; i < n; ++i) {
        const int tile = begin + i;
        mbar_wait(&x_empty[xs], xph ^ 1);
        mbar_arrive_expect_tx(&x_full[xs], kQ2Slab);
        tma_load_2d(xring + static_cast<size_t>(xs) * kQ2Slab, &mapX, &x_full[xs], 0, tile * kTileM);
        if (++xs == kQ2XStages) {
          xs = 0;
          xph ^= 1;
        }
        const int b = tile / p.tiles_per_image;
        mbar_wait(&w_empty[ws], wph ^ 1);
        mbar_arrive_expect_tx(&w_full[ws], kQ2Slab);
        for (int kc = 0; kc < 2; ++kc)
          tma_load_2d(wring + static_cast<size_t>(ws) * kQ2Slab + kc * (kQ2Slab / 2), &mapWeff, &w_full[ws], kc * 64,
                      b * p.c_pad);
        if (++ws == kQ2WStages) {
          ws = 0;
          wph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: G1(t0); then per tile i: G1(t_{i+1}), G2(t_i) =====================
    const uint32_t idesc1 = make_idesc_bf16(kTileM, 128);
    const uint32_t idesc2 = make_idesc_bf16(kTileM, 64);
    const uint64_t desc_fixed = make_sw128_desc(0);
    const uint32_t wring_lo = (smem_u32(wring) & 0x3FFFF) >> 4, wq_lo = (smem_u32(wq) & 0x3FFFF) >> 4,
                   a2_lo = (smem_u32(a2) & 0x3FFFF) >> 4, x_lo = (smem_u32(xring) & 0x3FFFF) >> 4, slab_lo = kQ2Slab >> 4;
    int ws = 0, xs = 0;
    uint32_t wph = 0, xph = 0;
    auto gemm1 = [&](int i) {
      const int t = i & 1;
      mbar_wait(&x_full[xs], xph);             // the raw tile is the operand: no need to wait for the statistics
      mbar_wait(&acc1_empty[t], ((i >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + t * kAccStride;
      const uint64_t adesc = desc_fixed | (x_lo + xs * slab_lo);
      const uint64_t bdesc = desc_fixed | wq_lo;
      if (elect_one()) {
        umma_bf16(d, adesc, bdesc, idesc1, 0u);
        umma_bf16(d, adesc + 2, bdesc + 2, idesc1, 1u);
        umma_bf16(d, adesc + 4, bdesc + 4, idesc1, 1u);
        umma_bf16(d, adesc + 6, bdesc + 6, idesc1, 1u);
        umma_commit(&acc1_full[t]);
      }
      __syncwarp();
      if (++xs == kQ2XStages) {
        xs = 0;
        xph ^= 1;
      }
    };
    auto gemm2 = [&](int i) {
      const int t = i & 1;
      mbar_wait(&a2_full[t], (i >> 1) & 1);      // q tile staged (and D2[t] drained: same groups, program order)
      mbar_wait(&w_full[ws], wph);
      tc_fence_after();
      const uint32_t d = tmem_base + t * kAccStride + 128;
      if (elect_one()) {
#pragma unroll
        for (int kc = 0; kc < 2; ++kc) {
          const uint64_t adesc = desc_fixed | (a2_lo + (t * 2 + kc) * slab_lo);
          const uint64_t bdesc = desc_fixed | (wring_lo + ws * slab_lo + kc * (slab_lo >> 1));
          umma_bf16(d, adesc, bdesc, idesc2, kc ? 1u : 0u);
          umma_bf16(d, adesc + 2, bdesc + 2, idesc2, 1u);
          umma_bf16(d, adesc + 4, bdesc + 4, idesc2, 1u);
          umma_bf16(d, adesc + 6, bdesc + 6, idesc2, 1u);
        }
        umma_commit(&w_empty[ws]);
        umma_commit(&d2_full[t]);
      }
      __syncwarp();
      if (++ws == kQ2WStages) {
        ws = 0;
        wph ^= 1;
      }
    };
    mbar_wait(wq_full, 0);
    if (n > 0) gemm1(0);
    for (int i = 0; i < n; ++i) {
      if (i + 1 < n) gemm1(i + 1);
      gemm2(i);
    }
  } else if (warp < 6) {
    // ===================== row statistics: thread = pixel row of the landed tile =====================
    const int row = (warp - 2) * 32 + lane;
    int xs = 0;
    uint32_t xph = 0;
    for (int i = 0; i < n; ++i) {
      mbar_wait(&x_full[xs], xph);
      float mean, var;
      row_moments64(smem_u32(xring) + xs * kQ2Slab, row, mean, var);
      rowa[xs * kTileM + row] = rsqrtf(var + p.prenorm_eps) * 1.4426950408889634f;
      mbar_arrive(&st_full[xs]);
      if (++xs == kQ2XStages) {
        xs = 0;
        xph ^= 1;
      }
    }
  } else {
    // ===================== epilogue: tile pair t, column half s, thread = pixel row =====================
    const int quad = warp & 3;
    const int eg = (warp - 6) >> 2;          // 0..3
    const int t = eg >> 1, s = eg & 1;
    const int row = quad * 32 + lane;
    const bool leader = (warp - 6 - 4 * eg) == 0 && lane == 0 && s == 0;   // one thread per tile pair
    griddep_wait();
    uint8_t* qt = a2 + t * 2 * kQ2Slab;                     // this pair's q tile; slab 0 is reused as the output staging tile
    const uint32_t qt_s = smem_u32(qt);
    float* st_mine = stat + ((t * 2 + s) * kTileM + row) * 2;
    const float* st_other = stat + ((t * 2 + (s ^ 1)) * kTileM + row) * 2;
    const uint32_t acc1 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + t * kAccStride;
    const uint32_t acc2 = acc1 + 128;
    const int bar_id = 1 + t;                               // named barrier of the 256 threads that share tile pair t
    const float sh0 = p.q_shift[2 * s], sh1 = p.q_shift[2 * s + 1];
    // second half of the tile: bias, output LayerNorm (this half's 32 channels, statistics shared with the other half),
    // gain, residual from the raw tile, bf16 staging.  A template over the half so that bias / gain are constant-bank
    // operands of the arithmetic instructions themselves.
    auto finish = [&](auto half, int xs) {
      constexpr int S = decltype(half)::value;
      uint32_t r[32];
      tmem_ld32(acc2 + 32 * S, r);
      tmem_ld_wait();
      uint64_t v[16];
      uint64_t sum2 = 0ull, sq2 = 0ull;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        v[j] = add_f32x2(pack_f32x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1])),
                         pack_f32x2(p.bias[32 * S + 2 * j], p.bias[32 * S + 2 * j + 1]));
        sum2 = add_f32x2(sum2, v[j]);
        sq2 = fma_f32x2(v[j], v[j], sq2);
      }
      float sa, sb, qa, qb;
      unpack_f32x2(sum2, sa, sb);
      unpack_f32x2(sq2, qa, qb);
      const float sum = sa + sb, sq = qa + qb;
      *reinterpret_cast<float2*>(st_mine) = make_float2(sum, sq);
      asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
      const float2 o = *reinterpret_cast<const float2*>(st_other);
      const float mean = (sum + o.x) * (1.0f / 64.0f);
      const float var = fmaxf((sq + o.y) * (1.0f / 64.0f) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + p.ln_eps);
      const uint64_t r2 = pack_f32x2(rstd, rstd), nm2 = pack_f32x2(-mean * rstd, -mean * rstd);
      // residual = the raw tile (this half's 64 bytes of the row: 16-byte pieces 4 S .. 4 S + 3, swizzled by the row)
      const uint32_t xrow = smem_u32(xring) + xs * kQ2Slab + row * 128;
      const uint32_t orow = qt_s + row * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t u[4];
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3])
                     : "r"(xrow + (((4 * S + q) ^ (row & 7)) << 4)));
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = 4 * q + e;
          const uint64_t res = pack_f32x2(__uint_as_float(u[e] << 16), __uint_as_float(u[e] & 0xffff0000u));
          const uint64_t tn = fma_f32x2(v[j], r2, nm2);                                     // (v - mean) rstd
          float lo, hi;
          unpack_f32x2(fma_f32x2(tn, pack_f32x2(p.ln_g[32 * S + 2 * j], p.ln_g[32 * S + 2 * j + 1]), res), lo, hi);
          w[e] = pack_bf16(lo, hi);
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(orow + (((4 * S + q) ^ (row & 7)) << 4)), "r"(w[0]),
                     "r"(w[1]), "r"(w[2]), "r"(w[3])
                     : "memory");
      }
    };
    for (int i = t; i < n; i += 2) {
      const int tile = begin + i;
      const uint32_t ph = (i >> 1) & 1;
      const int xs = i % kQ2XStages;
      const uint32_t xph = (i / kQ2XStages) & 1;
      // the previous output store of this pair must have finished reading the staging slab before q is written there
      if (leader) tma_store_wait_read();
      asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
      // ---- epilogue 1: q = softmax over the 32 channels of each of this half's two heads, * 32^-0.5 -> bf16 A operand
      mbar_wait(&acc1_full[t], ph);
      tc_fence_after();
      {
        uint32_t r[2][32];
        tmem_ld32(acc1 + 64 * s, r[0]);
        tmem_ld32(acc1 + 64 * s + 32, r[1]);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&acc1_empty[t]);           // the accumulator is in registers
        mbar_wait(&st_full[xs], xph);
        const float a = rowa[xs * kTileM + row];           // rstd log2e: q_pre log2e = acc a
        const uint64_t a2v = pack_f32x2(a, a);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float w[32];
          float ml = hh ? sh1 : sh0;                       // data-independent bound (>= the row maximum) ...
          if (p.use_max) {                                 // ... or the row maximum itself when the bound is too loose
            float m4[4] = {__uint_as_float(r[hh][0]), __uint_as_float(r[hh][1]), __uint_as_float(r[hh][2]),
                           __uint_as_float(r[hh][3])};
#pragma unroll
            for (int j = 4; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(r[hh][j]));
            ml = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * a;
          }
          const uint64_t mml = pack_f32x2(-ml, -ml);
          uint64_t sum2[2] = {0ull, 0ull};
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float a0, a1;
            unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[hh][j]), __uint_as_float(r[hh][j + 1])), a2v, mml), a0, a1);
            w[j] = ex2_approx(a0);
            w[j + 1] = ex2_approx(a1);
            sum2[(j >> 1) & 1] = add_f32x2(sum2[(j >> 1) & 1], pack_f32x2(w[j], w[j + 1]));
          }
          float s0, s1;
          unpack_f32x2(add_f32x2(sum2[0], sum2[1]), s0, s1);
          const float inv = __fdividef(0.17677669529663687f, s0 + s1);
          const uint64_t inv2 = pack_f32x2(inv, inv);
#pragma unroll
          for (int j = 0; j < 32; j += 2) unpack_f32x2(mul_f32x2(pack_f32x2(w[j], w[j + 1]), inv2), w[j], w[j + 1]);
          chunk_stage_bf16_s(qt_s, row, 64 * s + 32 * hh, w);
        }
      }
      fence_proxy_async();                                // generic-proxy smem writes -> visible to the tensor core
      mbar_arrive(&a2_full[t]);
      // ---- epilogue 2 ----
      mbar_wait(&d2_full[t], ph);                         // GEMM 2 done: D2 complete, the q tile is free again
      tc_fence_after();
      if (s == 0) finish(std::integral_constant<int, 0>{}, xs);
      else finish(std::integral_constant<int, 1>{}, xs);
      tc_fence_before();
      fence_proxy_async();
      asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
      if (leader) {
        mbar_arrive(&x_empty[xs]);                        // every thread of the pair has read its residual
        tma_store_2d(&mapOut, qt, 0, tile * kTileM);
        tma_store_commit();
      }
    }
    if (leader) tma_store_wait_read();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace dac

using namespace dac;

int dac_qout2_smem_bytes() {
  return (1 + kQ2XStages + kQ2WStages + 4) * (int)kQ2Slab + kQ2XStages * kTileM * (int)sizeof(float) +
         2 * 2 * kTileM * 2 * (int)sizeof(float) + 1024 + 512;
}

int dac_qout2_launch(const CUtensorMap& mapX, const CUtensorMap& mapWq, const CUtensorMap& mapWeff,
                     const CUtensorMap& mapOut, const Qout2Params& kp, int grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(linattn_qout2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dac_qout2_smem_bytes());
    if (e != cudaSuccess) return set_error(-12, "dac_linattn_qout: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  launch_k(linattn_qout2_kernel, dim3(grid), dim3(kQ2Threads), dac_qout2_smem_bytes(), st, mapX, mapWq, mapWeff, mapOut, kp);
  return check_launch("linattn_qout2_kernel");
}
