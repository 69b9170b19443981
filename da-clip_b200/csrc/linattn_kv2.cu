// LinearAttention, key side for 64-channel inputs, second generation (round 2): the values are never computed and the
// PreNorm is folded into the GEMMs - the kernel reads the RAW activation tensor once and writes a few KB per image.
// (module_util.py:89-97 PreNorm + :77-86 channel LayerNorm + :170-177 k-softmax / context of the reference.)
//
// 1. Values.  context[h][d][e] = sum_px softmax_px(k)[d] v[e] with v = W_v xn, so context = (P^T Xn) W_v^T / S: the kernel
//    accumulates G[(h,d)][c] = sum_px P[px][(h,d)] xn[px][c] and S[(h,d)] = sum_px P[px][(h,d)] per image and the fold kernel
//    (dac_linattn_fold_g, linattn.cu) applies the constant per-head matrix M_h = W_out,h W_v,h afterwards.  Against round
//    1's linattn_kv_kernel<64> per 128-pixel tile: GEMM 1 is N = 128 instead of 256 (k only), the epilogue handles 128
//    instead of 256 accumulator columns and stages P only, GEMM 2 is ONE N = 80 MMA per K step (the 64 channels of the
//    activation tile itself as the MN-major B operand + 16 more columns for S) instead of an N = 128 and an N = 16 one.
// 2. PreNorm.  xn = (x - mean) rstd per pixel (gain folded into the weights).  With row-centred weights W_c = W -
//    rowmean(W) the mean term vanishes (W_c 1 = 0): W xn = rstd (W_c x).  So GEMM 1 runs on the raw bf16 tile and the
//    epilogue multiplies by the pixel's rstd - inside the FFMA that feeds exp2, for free.  In GEMM 2 the pixel is the K
//    index, so rstd is folded into the A operand: P' = P rstd = exp2(k log2e - shift + log2 rstd), again inside the same
//    FFMA; G' = P'^T x then equals P^T (x rstd), and W_v is row-centred too (in M_h), which removes the mean there.  S
//    needs the unscaled P: the 16 extra B columns hold 1 / rstd of the pixel (bf16), so S = sum_px P' / rstd.
//    Four warps compute the row moments of each landed tile (8 LDS.128 + ~70 packed fp32 ops per row - no normalised
//    tile is written anywhere); round 1 ran `prenorm` as its own pass (56 us at 256^2, batch 16).
//
//   k_raw = W_kc x                                       GEMM 1 per 128-pixel tile: M128 x N128 x K64 (two TMEM stages)
//   P' = exp2(k_raw rstd log2e - c_h + log2 rstd)        epilogue: 64 columns (two heads) per thread, bf16 [pixel][channel]
//   G' | S += P'^T [x | 1 / rstd]                        GEMM 2: M128 x N80 x K128(pixels), both operands MN-major
// TMEM: [0,128) / [128,256) GEMM-1 accumulators, [256,336) / [384,464) {G', S} of alternating images.
// Roles (22 warps): warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 row statistics, warps 6-21 epilogue (two sets of
// eight on alternating tiles; a thread = one pixel row, two heads).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <new>

#include "../../include/dac_b200.h"
#include "common.h"
#include "linattn_kv_common.h"
#include "linattn_rowstats.cuh"
#include "tensormap.h"
#include "tile_common.cuh"

namespace dac {

// DAC_KV2_PROF (hand-built debug library only, tools/build_variant.py): per-warp cycle totals spent in each kind of wait
#ifdef DAC_KV2_PROF
#define KV2_T(slot, stmt)                                            \
  do {                                                               \
    const long long t0_ = clock64();                                 \
    stmt;                                                            \
    prof_acc[slot] += clock64() - t0_;                               \
  } while (0)
#else
#define KV2_T(slot, stmt) stmt
#endif

constexpr uint32_t kKv2Slab = kTileM * 128;     // 128 rows x 64 bf16 (16 KB)
constexpr uint32_t kKv2ColG = 256, kKv2GN = 80, kKv2GStride = 128;   // {G[64], S[16]} per image buffer
constexpr int kKv2Threads = 704;
constexpr int kKv2Stages = 6;
constexpr int kKv2InvSlabs = (kKv2Stages + 3) / 4;   // 1 / rstd columns: 16 of a slab's 64 columns per stage

__global__ void __launch_bounds__(kKv2Threads, 1)
linattn_kv2_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW,
                   const __grid_constant__ Kv2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wres = smem;                                         // W_kc: 128 rows x 128 B
  uint8_t* pbuf = wres + kKv2Slab;                              // [2 buffers][2 slabs]: P' tile, channels 0-63 | 64-127
  uint8_t* ring = pbuf + 4 * kKv2Slab;                          // [kKv2Stages] raw activation tiles
  uint8_t* invr = ring + kKv2Stages * kKv2Slab;                 // BEHIND the ring: stage s owns columns [16 (s & 3), +16) of slab s >> 2
  float2* rowp = reinterpret_cast<float2*>(invr + kKv2InvSlabs * kKv2Slab);   // [kKv2Stages][128] {rstd log2e, log2 rstd}
  uint64_t* bars = reinterpret_cast<uint64_t*>(rowp + kKv2Stages * kTileM);
  uint64_t* full = bars;                 // [8]  TMA landed
  uint64_t* stat = bars + 8;             // [8]  row statistics written (count 128)
  uint64_t* empty = bars + 16;           // [8]  GEMM 2 has read the tile
  uint64_t* acc_full = bars + 24;        // [2]
  uint64_t* acc_empty = bars + 26;       // [2]  count 256 (the eight warps of the tile's epilogue set)
  uint64_t* p_full = bars + 28;          // [2]  count 256
  uint64_t* p_free = bars + 30;          // [2]
  uint64_t* g_done = bars + 32;          // [2]  the image's {G', S} is complete
  uint64_t* g_flushed = bars + 34;       // [2]  count 128
  uint64_t* w_full = bars + 36;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 37);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int begin, end;
  tile_range(p.tiles, begin, end);
  const int n = end - begin;
#ifdef DAC_KV2_PROF
  long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long prof_t0 = clock64();
#endif

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapX);
    tma_prefetch_desc(&mapW);
    for (int s = 0; s < kKv2Stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&stat[s], 128);
      mbar_init(&empty[s], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&acc_full[g], 1);
      mbar_init(&acc_empty[g], 256);
      mbar_init(&p_full[g], 256);
      mbar_init(&p_free[g], 1);
      mbar_init(&g_done[g], 1);
      mbar_init(&g_flushed[g], 128);
    }
    mbar_init(w_full, 1);
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
  griddep_launch();   // see conv_kernel.cuh: set-up and the W_k load overlap the previous kernel's tail

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, kKv2Slab);
      tma_load_2d(wres, &mapW, w_full, 0, 0);
      griddep_wait();                        // the activation tensor is the previous kernel's output
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < n; ++i) {
        KV2_T(0, mbar_wait(&empty[stage], phase ^ 1));
        mbar_arrive_expect_tx(&full[stage], kKv2Slab);
        tma_load_2d(ring + static_cast<size_t>(stage) * kKv2Slab, &mapX, &full[stage], 0, (begin + ((p.dbg & 16) ? (i & 3) : i)) * kTileM);
        if (++stage == kKv2Stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: k(0), k(1); then per tile i: k(i+2), g(i) =====================
    const uint32_t idesc_k = make_idesc_bf16(kTileM, 128);
    const uint32_t idesc_g = make_idesc_bf16(kTileM, kKv2GN) | (1u << 15) | (1u << 16);   // A and B MN-major
    const uint64_t desc_k = make_sw128_desc(0);
    const uint64_t desc_nolbo = desc_k & ~(static_cast<uint64_t>(0x3FFF) << 16);
    // MN-major P' operand spanning two 64-channel swizzle atoms: leading byte offset = distance between the slabs
    const uint64_t desc_p = desc_nolbo | (static_cast<uint64_t>(kKv2Slab >> 4) << 16);
    const uint32_t ring_lo = (smem_u32(ring) & 0x3FFFF) >> 4, w_lo = (smem_u32(wres) & 0x3FFFF) >> 4,
                   p_lo0 = (smem_u32(pbuf) & 0x3FFFF) >> 4, invr_lo = (smem_u32(invr) & 0x3FFFF) >> 4,
                   slab_lo = kKv2Slab >> 4;
    int kstage = 0, gstage = 0;
    uint32_t kphase = 0, gphase = 0;
    int cur_img = -1, gb = 1;
    uint32_t gacc = 0;
    uint32_t nimg = 0;                       // images started by this CTA
    auto kgemm = [&](int i) {
      KV2_T(0, mbar_wait(&full[kstage], kphase));      // GEMM 1 reads the raw tile: no need to wait for the statistics
      KV2_T(1, mbar_wait(&acc_empty[i & 1], ((i >> 1) & 1) ^ 1));
      tc_fence_after();
      const uint64_t adesc = desc_k | (ring_lo + kstage * slab_lo);
      const uint64_t bdesc = desc_k | w_lo;
      const uint32_t d = tmem_base + (i & 1) * 128;
      if (elect_one()) {
        if (!(p.dbg & 32)) {
          umma_bf16(d, adesc, bdesc, idesc_k, 0u);
          umma_bf16(d, adesc + 2, bdesc + 2, idesc_k, 1u);
          umma_bf16(d, adesc + 4, bdesc + 4, idesc_k, 1u);
          umma_bf16(d, adesc + 6, bdesc + 6, idesc_k, 1u);
        }
        umma_commit(&acc_full[i & 1]);
      }
      __syncwarp();
      if (++kstage == kKv2Stages) {
        kstage = 0;
        kphase ^= 1;
      }
    };
    auto ggemm = [&](int i) {
      const int img = (begin + i) / p.tiles_per_image;
      if (img != cur_img) {
        if (cur_img >= 0) {                  // the finished image's {G', S} goes to the flushing group
          if (elect_one()) umma_commit(&g_done[gb]);
          __syncwarp();
        }
        cur_img = img;
        gb ^= 1;
        if (nimg >= 2) {                     // the accumulator's previous image must have been read out
          KV2_T(4, mbar_wait(&g_flushed[gb], ((nimg >> 1) - 1) & 1));
          tc_fence_after();
        }
        ++nimg;
        gacc = 0;
      }
      KV2_T(2, mbar_wait(&stat[gstage], gphase));      // the 1 / rstd columns of this stage (the epilogue waited for them too)
      KV2_T(3, mbar_wait(&p_full[i & 1], (i >> 1) & 1));
      tc_fence_after();
      const uint32_t a_lo = p_lo0 + (i & 1) * 2 * slab_lo;
      const uint32_t x_lo = ring_lo + gstage * slab_lo;
      // B = [raw tile | 1 / rstd]: the second 64-column atom (16 columns used) starts at this stage's columns of its
      // 1 / rstd slab; the start offset of 32 B per stage selects them (the swizzle acts on the final address bits)
      const uint32_t i_lo = invr_lo + (gstage >> 2) * slab_lo + (gstage & 3) * 2;
      const uint64_t desc_b = desc_nolbo | (static_cast<uint64_t>(i_lo - x_lo) << 16);
      const uint32_t d = tmem_base + kKv2ColG + gb * kKv2GStride;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8 && !(p.dbg & 2); ++ks)       // 16 pixels (rows of the [pixel][channel] tiles) per K step: +2048 B
          umma_bf16(d, desc_p | (a_lo + ks * 128), desc_b | (x_lo + ks * 128), idesc_g, gacc | (ks ? 1u : 0u));
        umma_commit(&p_free[i & 1]);
        umma_commit(&empty[gstage]);
      }
      __syncwarp();
      gacc = 1;
      if (++gstage == kKv2Stages) {
        gstage = 0;
        gphase ^= 1;
      }
    };
    mbar_wait(w_full, 0);
    if (n > 0) kgemm(0);
    if (n > 1) kgemm(1);
    for (int i = 0; i < n; ++i) {
      // GEMM 1 of tile i+2 goes out as soon as tile i's accumulator has been read - early in tile i's epilogue - so that
      // it runs under those exponentials; GEMM 2 of tile i needs the END of that epilogue.  (The other order chains
      // accumulator -> epilogue -> GEMM 2 -> GEMM 1 -> accumulator serially per epilogue set: measured 102 vs 82 us.)
      if (i + 2 < n) kgemm(i + 2);
      ggemm(i);
    }
    if (n > 0) {
      if (elect_one()) umma_commit(&g_done[gb]);
      __syncwarp();
    }
  } else if (warp < 6) {
    // ===================== row statistics: thread = pixel row of the tile =====================
    const int row = (warp - 2) * 32 + lane;
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < n; ++i) {
      KV2_T(0, mbar_wait(&full[stage], phase));
      // (the stage's previous tile is done with its 1 / rstd columns and rowp entries: the producer waited for `empty`)
      float mean = 0.f, var = 1.f;
      if (!(p.dbg & 8)) row_moments64(smem_u32(ring) + stage * kKv2Slab, row, mean, var);
      const float ve = var + p.ln_eps;
      const float rstd = rsqrtf(ve);
      rowp[stage * kTileM + row] = make_float2(rstd * 1.4426950408889634f, -0.5f * __log2f(ve));
      const uint32_t iv = pack_bf16(ve * rstd, ve * rstd);       // 1 / rstd = sqrt(var + eps)
      const uint32_t ibase = smem_u32(invr) + (stage >> 2) * kKv2Slab + row * 128;
#pragma unroll
      for (int q = 0; q < 2; ++q)
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(ibase + (((2 * (stage & 3) + q) ^ (row & 7)) << 4)), "r"(iv)
                     : "memory");
      fence_proxy_async();                   // generic-proxy writes -> visible to the tensor core
      mbar_arrive(&stat[stage]);
      if (++stage == kKv2Stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue: set e = tiles i = e (mod 2), group j = heads 2j, 2j+1, thread = pixel row =====================
    // Two sets of eight warps on alternating tiles (own accumulator stage, own P' buffer), so that the latencies of one
    // tile's chain - accumulator ready, tcgen05.ld, exponentials, staging, proxy fence, GEMM 2 - overlap the other's.
    const int quad = warp & 3;
    const int e = (warp - 6) >> 3;           // 0..1
    const int j = ((warp - 6) >> 2) & 1;     // 0..1
    const int row = quad * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    // The shift only has to keep exp() in range: softmax over the pixels of channel d is invariant to ANY per-channel
    // constant (G and S carry the same factor), so ONE scalar - the largest bound of the head - serves all 32 channels.
    const float sh[2] = {p.shift_max[2 * j], p.shift_max[2 * j + 1]};
    const bool flusher = e == 0 && j == 0;   // warps 6-9: TMEM lane quarter `quad` = head `quad` of {G', S}
    uint32_t nflush = 0;
    auto flush = [&](int img) {
      const int gbuf = nflush & 1;
      mbar_wait(&g_done[gbuf], (nflush >> 1) & 1);
      tc_fence_after();
      ++nflush;
      const int slot = static_cast<int>(blockIdx.x) - tile_owner(img * p.tiles_per_image, p.tiles, gridDim.x);
      float* rec = p.ctx_acc + ((static_cast<long long>(img) * 4 + quad) * p.slots + slot) * kKvGRec;
      const uint32_t g = lane_base + kKv2ColG + gbuf * kKv2GStride;
      float v[32];
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        chunk_from_tmem(g + c, v);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(rec + lane * 64 + c + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
      uint32_t r[16];
      tmem_ld16(g + 64, r);
      tmem_ld_wait();
      rec[2048 + lane] = __uint_as_float(r[0]);
      tc_fence_before();
      mbar_arrive(&g_flushed[gbuf]);
    };
    if (flusher) griddep_wait();             // ctx_acc may still be read by the previous step's fold kernel
    int next_flush = n > 0 ? begin / p.tiles_per_image : 0;       // images of this CTA not yet written out (flusher only)
    const uint32_t pt = smem_u32(pbuf) + static_cast<uint32_t>(e) * 2 * kKv2Slab;
    for (int i = e; i < n; i += 2) {
      const uint32_t ph = (i >> 1) & 1;
      const int stage = i % kKv2Stages;
      const uint32_t sphase = (i / kKv2Stages) & 1;
      KV2_T(0, mbar_wait(&acc_full[e], ph));
      tc_fence_after();
      uint32_t rk2[2][32];
      KV2_T(1, tmem_ld32(lane_base + e * 128 + 64 * j, rk2[0]); tmem_ld32(lane_base + e * 128 + 64 * j + 32, rk2[1]);
            tmem_ld_wait());
      tc_fence_before();
      mbar_arrive(&acc_empty[e]);                                            // the accumulator is in registers
      KV2_T(2, mbar_wait(&stat[stage], sphase));
      const float2 rp = rowp[stage * kTileM + row];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const uint32_t(&rk)[32] = rk2[hh];
        const uint64_t a2 = pack_f32x2(rp.x, rp.x), c2 = pack_f32x2(rp.y - sh[hh], rp.y - sh[hh]);
        float v[32];
#ifdef DAC_KV2_PROF
        const long long te0 = clock64();
#endif
#pragma unroll
        for (int q = 0; q < 32; q += 2) {
          float a0, a1;
          unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(rk[q]), __uint_as_float(rk[q + 1])), a2, c2), a0, a1);
          v[q] = (p.dbg & 1) ? a0 : ex2_approx(a0);
          v[q + 1] = (p.dbg & 1) ? a1 : ex2_approx(a1);
        }
#ifdef DAC_KV2_PROF
        prof_acc[4] += clock64() - te0;
#endif
        if (hh == 0) KV2_T(3, mbar_wait(&p_free[e], ph ^ 1));                // GEMM 2 of the buffer's previous tile is done
        KV2_T(5, if (!(p.dbg & 4)) chunk_stage_bf16_s(pt, row, 64 * j + 32 * hh, v));
      }
      KV2_T(6, fence_proxy_async(); mbar_arrive(&p_full[e]));
      if (flusher) {
        const int img = (begin + i) / p.tiles_per_image;
        while (next_flush < img) flush(next_flush++);                        // complete since GEMM 2 moved on to `img`
      }
    }
    if (flusher && n > 0) {
      const int last = (end - 1) / p.tiles_per_image;
      while (next_flush <= last) flush(next_flush++);
    }
  }

#ifdef DAC_KV2_PROF
  if (blockIdx.x == 0 && lane == 0 && p.prof) {
    prof_acc[7] = clock64() - prof_t0;
    for (int i = 0; i < 8; ++i) p.prof[warp * 8 + i] = prof_acc[i];
  }
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace dac

using namespace dac;

int dac_kv2_smem_bytes() {
  return (1 + 4 + kKv2Stages + kKv2InvSlabs) * (int)kKv2Slab + kKv2Stages * kTileM * (int)sizeof(float2) + 1024 + 512;
}

int dac_kv2_launch(const CUtensorMap& mapX, const CUtensorMap& mapW, const Kv2Params& kp, int grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(linattn_kv2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dac_kv2_smem_bytes());
    if (e != cudaSuccess) return set_error(-12, "dac_linattn_kv: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
#ifdef DAC_KV2_PROF
  static long long* prof_buf = nullptr;
  if (!prof_buf) cudaMallocManaged(&prof_buf, 22 * 8 * sizeof(long long));
  if (getenv("DAC_KV2_PROF_DUMP")) {        // print the totals of the PREVIOUS launch
    cudaDeviceSynchronize();
    for (int w = 0; w < 22; ++w) {
      printf("warp %2d:", w);
      for (int i = 0; i < 8; ++i) printf(" %9lld", prof_buf[w * 8 + i]);
      printf("\n");
    }
  }
  Kv2Params kq = kp;
  kq.prof = prof_buf;
  linattn_kv2_kernel<<<grid, kKv2Threads, dac_kv2_smem_bytes(), st>>>(mapX, mapW, kq);
#else
  launch_k(linattn_kv2_kernel, dim3(grid), dim3(kKv2Threads), dac_kv2_smem_bytes(), st, mapX, mapW, kp);
#endif
  return check_launch("linattn_kv2_kernel");
}
