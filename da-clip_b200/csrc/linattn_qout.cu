// LinearAttention, query side, as ONE chained-GEMM kernel (module_util.py:170-185 of the reference):
//
//   q    = softmax_channels-of-head(W_q xn) * 32^-0.5            GEMM 1: [128 px, C] x [C, 128]
//   out  = LayerNorm_c(W_eff[b] q + bias) * g + x                GEMM 2: [128 px, 128] x [128, C]
//
// where W_eff[b] = W_out . ctx[b] is the per-image weight dac_linattn_fold built from the k|v context.  The softmaxed
// q tile never leaves the SM: the epilogue of GEMM 1 writes it (bf16, 128B-swizzled K-major - the layout TMA would
// have produced) into shared memory, where it is the A operand of GEMM 2.  Per 128-pixel tile the kernel reads the
// normalised input and the residual once and writes the output once; the unfused pair (to_q + to_out) also wrote and
// re-read the [pixels, 128] q tensor (2 x 268 MB per level-0 launch at batch 16).
//
// Roles (320 threads, one persistent CTA per SM): warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warps 2-9 = two
// epilogue groups, group g owning every second tile of the CTA, its accumulators (TMEM columns [256 g, 256 g + 128)
// for GEMM 1 and [256 g + 128, 256 g + 128 + C) for GEMM 2) and its q tile.  The issuer runs GEMM 1 of tile i+1
// BEFORE GEMM 2 of tile i, so the tensor pipe has work while group (i & 1) computes the softmax of tile i.
#include <cuda.h>
#include <cuda_runtime.h>
#include <new>

#include "../../include/dac_b200.h"
#include "common.h"
#include "linattn_qout_common.h"
#include "tile_common.cuh"
#include "tensormap.h"

namespace dac {

constexpr uint32_t kSlab = kTileM * 128;   // 128 rows x 64 bf16: one swizzled K chunk of a 128-row operand (16 KB)

struct QoutParams {
  int tiles, tiles_per_image, c_pad, stages;
  const float* bias;
  const float* ln_g;
  float ln_eps;
  const float* ln_stats;   // folded PreNorm: per-pixel {mean, rstd} of the RAW input row (NULL: the input is normalised)
  const float* ln_colsum;  // [128] sum_c W'_q[n][c] of the bf16 weight rows
};

template <int C>
__global__ void __launch_bounds__(kThreads, 1)
linattn_qout_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapWq,
                    const __grid_constant__ CUtensorMap mapWeff, const __grid_constant__ CUtensorMap mapOut,
                    const __grid_constant__ CUtensorMap mapRes, const __grid_constant__ QoutParams p) {
  constexpr int kCh = C / 64;                       // K chunks of GEMM 1 = output slabs of GEMM 2
  constexpr uint32_t kB2Bytes = C * 128;            // one K chunk of W_eff[b]: C rows x 128 B
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wq = smem;                                              // [kCh] slabs, resident
  uint8_t* ring = wq + kCh * kSlab;                                // [stages] x 16 KB
  uint8_t* a2 = ring + static_cast<size_t>(p.stages) * kSlab;      // [2 groups][2 slabs]: q tile, then output staging
  uint8_t* rbuf = a2 + 4 * kSlab;                                  // C == 64: [2 groups] residual landing tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(rbuf + (C == 64 ? 2 * kSlab : 0));
  uint64_t* full = bars;
  uint64_t* empty = bars + kMaxStages;
  uint64_t* acc1_full = bars + 2 * kMaxStages;
  uint64_t* acc1_empty = acc1_full + 2;
  uint64_t* a2_full = acc1_full + 4;
  uint64_t* d2_full = acc1_full + 6;
  uint64_t* res_bar = acc1_full + 8;
  uint64_t* wq_full = acc1_full + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc1_full + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int begin, end;
  tile_range(p.tiles, begin, end);
  const int n = end - begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapX);
    tma_prefetch_desc(&mapWq);
    tma_prefetch_desc(&mapWeff);
    tma_prefetch_desc(&mapOut);
    tma_prefetch_desc(&mapRes);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&acc1_full[g], 1);
      mbar_init(&acc1_empty[g], 1);
      mbar_init(&a2_full[g], 1);
      mbar_init(&d2_full[g], 1);
      mbar_init(&res_bar[g], 1);
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
    // ===================== TMA producer: A1(t0); then per tile i: A1(t_{i+1}), B2(t_i) =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(wq_full, kCh * kSlab);
      for (int ck = 0; ck < kCh; ++ck) tma_load_2d(wq + ck * kSlab, &mapWq, wq_full, ck * 64, 0);
      griddep_wait();
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() {
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      };
      auto load_a1 = [&](int tile) {
        for (int ck = 0; ck < kCh; ++ck) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], kSlab);
          tma_load_2d(ring + static_cast<size_t>(stage) * kSlab, &mapX, &full[stage], ck * 64, tile * kTileM);
          advance();
        }
      };
      if (n > 0) load_a1(begin);
      for (int i = 0; i < n; ++i) {
        if (i + 1 < n) load_a1(begin + i + 1);
        const int b = (begin + i) / p.tiles_per_image;
        for (int kc = 0; kc < 2; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], kB2Bytes);
          tma_load_2d(ring + static_cast<size_t>(stage) * kSlab, &mapWeff, &full[stage], kc * 64, b * p.c_pad);
          advance();
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: G1(t0); then per tile i: G1(t_{i+1}), G2(t_i) =====================
    const uint32_t idesc1 = make_idesc_bf16(kTileM, 128);
    const uint32_t idesc2 = make_idesc_bf16(kTileM, C);
    const uint64_t desc_fixed = make_sw128_desc(0);
    const uint32_t ring_lo = (smem_u32(ring) & 0x3FFFF) >> 4, wq_lo = (smem_u32(wq) & 0x3FFFF) >> 4,
                   a2_lo = (smem_u32(a2) & 0x3FFFF) >> 4, slab_lo = kSlab >> 4;
    int stage = 0;
    uint32_t phase = 0;
    auto advance = [&]() {
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    };
    auto gemm1 = [&](int i) {
      const int g = i & 1;
      mbar_wait(&acc1_empty[g], ((i >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + g * kAccStride;
      for (int ck = 0; ck < kCh; ++ck) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t adesc = desc_fixed | (ring_lo + stage * slab_lo);
        const uint64_t bdesc = desc_fixed | (wq_lo + ck * slab_lo);
        if (elect_one()) {
          umma_bf16(d, adesc, bdesc, idesc1, ck ? 1u : 0u);
          umma_bf16(d, adesc + 2, bdesc + 2, idesc1, 1u);
          umma_bf16(d, adesc + 4, bdesc + 4, idesc1, 1u);
          umma_bf16(d, adesc + 6, bdesc + 6, idesc1, 1u);
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        advance();
      }
      if (elect_one()) umma_commit(&acc1_full[g]);
      __syncwarp();
    };
    auto gemm2 = [&](int i) {
      const int g = i & 1;
      mbar_wait(&a2_full[g], (i >> 1) & 1);      // q tile staged (and D2[g] drained: same group, program order)
      tc_fence_after();
      const uint32_t d = tmem_base + g * kAccStride + 128;
      for (int kc = 0; kc < 2; ++kc) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t adesc = desc_fixed | (a2_lo + (g * 2 + kc) * slab_lo);
        const uint64_t bdesc = desc_fixed | (ring_lo + stage * slab_lo);
        if (elect_one()) {
          umma_bf16(d, adesc, bdesc, idesc2, kc ? 1u : 0u);
          umma_bf16(d, adesc + 2, bdesc + 2, idesc2, 1u);
          umma_bf16(d, adesc + 4, bdesc + 4, idesc2, 1u);
          umma_bf16(d, adesc + 6, bdesc + 6, idesc2, 1u);
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        advance();
      }
      if (elect_one()) umma_commit(&d2_full[g]);
      __syncwarp();
    };
    mbar_wait(wq_full, 0);
    if (n > 0) gemm1(0);
    for (int i = 0; i < n; ++i) {
      if (i + 1 < n) gemm1(i + 1);
      gemm2(i);
    }
  } else {
    // ===================== epilogue groups =====================
    const int quad = warp & 3;
    const int group = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int gthread = threadIdx.x - 64 - group * 128;
    griddep_wait();
    uint8_t* qt = a2 + group * 2 * kSlab;                 // this group's q tile / output staging
    uint8_t* rt = (C == 64) ? rbuf + group * kSlab : qt;  // where the residual tile lands
    const uint32_t acc1 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + group * kAccStride;
    const uint32_t acc2 = acc1 + 128;
    float v[32];
    for (int i = group; i < n; i += 2) {
      const int tile = begin + i;
      const uint32_t ph = (i >> 1) & 1;
      if (C == 64) {
        // the output of a tile is staged IN the residual tile (rt) and stored from there, so nobody has to wait for the
        // previous store before writing the q tile; only the prefetch of the next residual does (one thread)
        if (gthread == 0) {
          tma_store_wait_read();
          mbar_arrive_expect_tx(&res_bar[group], kSlab);
          tma_load_2d(rt, &mapRes, &res_bar[group], 0, tile * kTileM);
        }
      } else {
        if (gthread == 0) tma_store_wait_read();          // the previous output store has finished reading qt
        asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
      }
      // ---- epilogue 1: q = softmax over the 32 channels of each head, * 32^-0.5 -> bf16 A operand of GEMM 2
      // folded PreNorm: q_raw = rstd * (W'_q x - mean * colsum); rstd > 0, so it only scales the softmax exponent
      float qa = 1.4426950408889634f, qmean = 0.f;
      if (p.ln_stats) {
        const float2 ms = __ldg(reinterpret_cast<const float2*>(p.ln_stats) + static_cast<long long>(tile) * kTileM + row);
        qmean = ms.x;
        qa = ms.y * 1.4426950408889634f;
      }
      mbar_wait(&acc1_full[group], ph);
      tc_fence_after();
      // two heads at a time: both TMEM loads issued before one wait so that the two softmax chains (max tree, exponentials,
      // sum tree, scale, pack) interleave in the instruction stream - this epilogue is a latency chain, not a throughput
      // problem (two warps per scheduler)
#pragma unroll
      for (int c = 0; c < 128; c += 64) {
        uint32_t r[2][32];
        tmem_ld32(acc1 + c, r[0]);
        tmem_ld32(acc1 + c + 32, r[1]);
        tmem_ld_wait();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float w[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) w[j] = __uint_as_float(r[hh][j]);
          if (p.ln_stats) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 cs = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + c + 32 * hh) + q);
              w[4 * q] = fmaf(-qmean, cs.x, w[4 * q]);
              w[4 * q + 1] = fmaf(-qmean, cs.y, w[4 * q + 1]);
              w[4 * q + 2] = fmaf(-qmean, cs.z, w[4 * q + 2]);
              w[4 * q + 3] = fmaf(-qmean, cs.w, w[4 * q + 3]);
            }
          }
          float m4[4] = {w[0], w[1], w[2], w[3]};
#pragma unroll
          for (int j = 4; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], w[j]);
          const float ml = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * qa;
          float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            w[j] = ex2_approx(fmaf(w[j], qa, -ml));
            s4[j & 3] += w[j];
          }
          const float inv = __fdividef(0.17677669529663687f, (s4[0] + s4[1]) + (s4[2] + s4[3]));
#pragma unroll
          for (int j = 0; j < 32; ++j) w[j] *= inv;
          chunk_stage_bf16(qt, row, c + 32 * hh, w);
        }
      }
      tc_fence_before();
      fence_proxy_async();                                // generic-proxy smem writes -> visible to the tensor core
      asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
      if (gthread == 0) {
        mbar_arrive(&a2_full[group]);
        mbar_arrive(&acc1_empty[group]);
      }
      // ---- epilogue 2: LayerNorm over the C channels of W_eff q + bias, gain, + residual
      mbar_wait(&d2_full[group], ph);                     // GEMM 2 done: D2 complete, qt free again
      tc_fence_after();
      if (C != 64 && gthread == 0) {
        mbar_arrive_expect_tx(&res_bar[group], kCh * kSlab);
        for (int s_ = 0; s_ < kCh; ++s_) tma_load_2d(rt + s_ * kSlab, &mapRes, &res_bar[group], s_ * 64, tile * kTileM);
      }
      float sum = 0.f;
      if (C == 64) {
        float w[2][32];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          uint32_t r[32];
          tmem_ld32(acc2 + k * 32, r);
#pragma unroll
          for (int j = 0; j < 32; ++j) w[k][j] = __uint_as_float(r[j]);
        }
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (p.bias) chunk_add_f32(p.bias + k * 32, w[k]);
#pragma unroll
          for (int j = 0; j < 32; ++j) sum += w[k][j];
        }
        const float mean = sum / C;
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            w[k][j] -= mean;
            ss = fmaf(w[k][j], w[k][j], ss);
          }
        const float rstd = rsqrtf(ss / C + p.ln_eps);
        mbar_wait(&res_bar[group], ph);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(p.ln_g + k * 32) + q);
            w[k][4 * q] *= rstd * g.x;
            w[k][4 * q + 1] *= rstd * g.y;
            w[k][4 * q + 2] *= rstd * g.z;
            w[k][4 * q + 3] *= rstd * g.w;
          }
          chunk_add_staged(rt, row, k * 32, w[k]);
          chunk_stage_bf16(rt, row, k * 32, w[k]);        // in place: same thread, same 16-byte pieces
        }
      } else {
        for (int c = 0; c < C; c += 32) {
          chunk_from_tmem(acc2 + c, v);
          if (p.bias) chunk_add_f32(p.bias + c, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) sum += v[j];
        }
        const float mean = sum / C;
        float ss = 0.f;
        for (int c = 0; c < C; c += 32) {
          chunk_from_tmem(acc2 + c, v);
          if (p.bias) chunk_add_f32(p.bias + c, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float d = v[j] - mean;
            ss = fmaf(d, d, ss);
          }
        }
        const float rstd = rsqrtf(ss / C + p.ln_eps);
        mbar_wait(&res_bar[group], ph);
        for (int c = 0; c < C; c += 32) {
          chunk_from_tmem(acc2 + c, v);
          if (p.bias) chunk_add_f32(p.bias + c, v);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(p.ln_g + c) + q);
            v[4 * q] = (v[4 * q] - mean) * rstd * g.x;
            v[4 * q + 1] = (v[4 * q + 1] - mean) * rstd * g.y;
            v[4 * q + 2] = (v[4 * q + 2] - mean) * rstd * g.z;
            v[4 * q + 3] = (v[4 * q + 3] - mean) * rstd * g.w;
          }
          chunk_add_staged(rt, row, c, v);                // residual landed in the staging tile itself
          chunk_stage_bf16(qt, row, c, v);
        }
      }
      tc_fence_before();
      fence_proxy_async();
      asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
      if (gthread == 0) {
        for (int s_ = 0; s_ < kCh; ++s_) tma_store_2d(&mapOut, (C == 64 ? rt : qt) + s_ * kSlab, s_ * 64, tile * kTileM);
        tma_store_commit();
      }
    }
    if (gthread == 0) tma_store_wait_read();
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

struct dac_qout_plan {
  CUtensorMap mapX, mapWq, mapWeff, mapOut, mapRes;
  QoutParams kp;
  Qout2Params kp2;
  int C, grid, smem;
  int prenorm;             // the input is the RAW tensor: linattn_qout2_kernel normalises its rows in shared memory
};

static int encode_2d(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t rows, uint64_t pitch_bytes,
                     uint32_t box_rows, const char* what) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return set_error(-10, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(-11, "cuTensorMapEncodeTiled(%s) failed: CUresult %d", what, (int)r);
  return 0;
}

extern "C" int dac_linattn_qout_create(const void* xn, const void* wq, const void* weff, int32_t c_pad,
                                       const void* res, void* out, const float* bias, const float* ln_g,
                                       float ln_eps, const float* ln_stats, const float* ln_colsum, int32_t B,
                                       int32_t hw, int32_t C, int32_t prenorm, float prenorm_eps, const float* q_shift,
                                       dac_qout_t* plan) {
  if (!xn || !wq || !weff || !res || !out || !ln_g || !plan) return set_error(-1, "dac_linattn_qout_create: null argument");
  *plan = nullptr;
  if (C != 64 && C != 128) return set_error(-2, "dac_linattn_qout_create: C must be 64 or 128 (got %d)", C);
  if (B <= 0 || hw <= 0 || hw % kTileM) return set_error(-2, "dac_linattn_qout_create: hw must be a multiple of 128");
  if (prenorm && (C != 64 || ln_stats || res != xn))
    return set_error(-2, "dac_linattn_qout_create: in-kernel PreNorm needs C = 64, no ln_stats and res == xn (the raw tensor)");
  if (q_shift && !prenorm) return set_error(-2, "dac_linattn_qout_create: q_shift belongs to the in-kernel PreNorm mode");
  if (c_pad < C || (c_pad & 7)) return set_error(-2, "dac_linattn_qout_create: bad c_pad");
  if ((reinterpret_cast<uintptr_t>(xn) | reinterpret_cast<uintptr_t>(wq) | reinterpret_cast<uintptr_t>(weff) |
       reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(bias) |
       reinterpret_cast<uintptr_t>(ln_g)) & 15)
    return set_error(-2, "dac_linattn_qout_create: pointers must be 16-byte aligned");
  dac_qout_plan* pl = new (std::nothrow) dac_qout_plan();
  if (!pl) return set_error(-3, "out of host memory");
  const uint64_t rows = static_cast<uint64_t>(B) * hw;
  int rc = encode_2d(&pl->mapX, xn, C, rows, C * 2ull, kTileM, "xn");
  if (!rc) rc = encode_2d(&pl->mapWq, wq, C, 128, C * 2ull, 128, "wq");
  if (!rc) rc = encode_2d(&pl->mapWeff, weff, 128, static_cast<uint64_t>(B) * c_pad, 256, C, "weff");
  if (!rc) rc = encode_2d(&pl->mapOut, out, C, rows, C * 2ull, kTileM, "out");
  if (!rc) rc = encode_2d(&pl->mapRes, res, C, rows, C * 2ull, kTileM, "res");
  if (rc) { delete pl; return rc; }
  QoutParams& k = pl->kp;
  k.tiles = static_cast<int>(rows / kTileM);
  k.tiles_per_image = hw / kTileM;
  k.c_pad = c_pad;
  k.bias = bias; k.ln_g = ln_g; k.ln_eps = ln_eps;
  k.ln_stats = ln_stats; k.ln_colsum = ln_colsum;
  if ((ln_stats != nullptr) != (ln_colsum != nullptr) ||
      ((reinterpret_cast<uintptr_t>(ln_stats) | reinterpret_cast<uintptr_t>(ln_colsum)) & 15)) {
    delete pl;
    return set_error(-2, "dac_linattn_qout_create: ln_stats and ln_colsum come together, 16-byte aligned");
  }
  const int fixed = (C / 64) * (int)kSlab + 4 * (int)kSlab + (C == 64 ? 2 * (int)kSlab : 0) + 1024 + 512;
  int stages = (227 * 1024 - fixed) / (int)kSlab;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 3) { delete pl; return set_error(-2, "dac_linattn_qout_create: does not fit shared memory"); }
  k.stages = stages;
  pl->smem = fixed + stages * (int)kSlab;
  pl->C = C;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  pl->grid = k.tiles < sms ? k.tiles : sms;
  cudaError_t e = C == 64 ? cudaFuncSetAttribute(linattn_qout_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, pl->smem)
                          : cudaFuncSetAttribute(linattn_qout_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, pl->smem);
  if (e != cudaSuccess) {
    delete pl;
    return set_error(-12, "dac_linattn_qout_create: cudaFuncSetAttribute(%d B smem): %s", pl->smem, cudaGetErrorString(e));
  }
  pl->prenorm = prenorm ? 1 : 0;
  if (prenorm) {
    Qout2Params& k2 = pl->kp2;
    k2.tiles = k.tiles;
    k2.tiles_per_image = k.tiles_per_image;
    k2.c_pad = c_pad;
    k2.ln_eps = ln_eps;
    k2.prenorm_eps = prenorm_eps;
    // bias, gain and the softmax bounds are constants of the layer: they travel in the kernel parameters
    cudaError_t ce = cudaMemcpy(k2.ln_g, ln_g, sizeof(k2.ln_g), cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess && bias) ce = cudaMemcpy(k2.bias, bias, sizeof(k2.bias), cudaMemcpyDeviceToHost);
    if (!bias) for (float& b : k2.bias) b = 0.f;
    k2.use_max = q_shift ? 0 : 1;
    for (float& q : k2.q_shift) q = 0.f;
    if (ce == cudaSuccess && q_shift) {
      float sh[128];
      ce = cudaMemcpy(sh, q_shift, sizeof(sh), cudaMemcpyDeviceToHost);
      for (int h = 0; h < 4 && ce == cudaSuccess; ++h) {
        float m = sh[h * 32];
        for (int d = 1; d < 32; ++d) m = sh[h * 32 + d] > m ? sh[h * 32 + d] : m;
        k2.q_shift[h] = m;
      }
    }
    if (ce != cudaSuccess) {
      delete pl;
      return set_error(-20, "dac_linattn_qout_create: reading the layer constants: %s", cudaGetErrorString(ce));
    }
  }
  *plan = pl;
  return 0;
}

extern "C" int dac_linattn_qout_launch(dac_qout_t pl, dac_stream_t stream) {
  if (!pl) return set_error(-1, "dac_linattn_qout_launch: null plan");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (pl->prenorm) return dac_qout2_launch(pl->mapX, pl->mapWq, pl->mapWeff, pl->mapOut, pl->kp2, pl->grid, st);
  if (pl->C == 64)
    launch_k(linattn_qout_kernel<64>, dim3(pl->grid), dim3(kThreads), pl->smem, st, pl->mapX, pl->mapWq, pl->mapWeff,
             pl->mapOut, pl->mapRes, pl->kp);
  else
    launch_k(linattn_qout_kernel<128>, dim3(pl->grid), dim3(kThreads), pl->smem, st, pl->mapX, pl->mapWq, pl->mapWeff,
             pl->mapOut, pl->mapRes, pl->kp);
  return check_launch("linattn_qout_kernel");
}

extern "C" void dac_linattn_qout_destroy(dac_qout_t pl) { delete pl; }
