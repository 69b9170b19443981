// LinearAttention, key/value side, entirely on the tcgen05 tensor cores (module_util.py:170-177 of the reference):
//
//   k | v = W_kv xn                      GEMM 1 per 128-pixel tile: two M128 x N128 MMAs (one per pair of heads)
//   P = exp(k - c_d)                     epilogue, data-independent shift c_d >= |k_d| (see DAC_EPI_KVCTX)
//   C[(h,d)][(h',e)] += P^T V            GEMM 2: M128 x N128 x K128(pixels), P and V as MN-major operands straight from
//   S[(h,d)]         += P^T 1            the [pixel][channel] tiles the epilogue wrote; N16 against a tile of ones
//
// k, v and P never leave the SM, and the context accumulates in TENSOR MEMORY across all the tiles of an image (the
// diagonal 32 x 32 blocks h = h' of C are the four per-head contexts; the off-diagonal blocks are the price of using
// the 128-wide MMA and cost nothing that matters: the kernel is bound by its epilogue).  At an image boundary and at the
// end the accumulator is read once and added into ctx_acc (the record format of dac_linattn_fold, nchunks = 1).
//
// Roles (320 threads, one persistent CTA per SM): warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 / 6-9 epilogue
// groups.  Both groups work on EVERY tile: group g owns heads 2g, 2g+1 (weight rows are packed k_2g k_2g+1 v_2g v_2g+1
// per group, so its N128 accumulator holds exactly its columns) and the 64-channel P / V slabs of those heads.
// TMEM: [0,128) / [128,256) GEMM-1 accumulators of group 0 / 1, [256,384) C, [384,400) S.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdlib>
#include <new>

#include "../../include/dac_b200.h"
#include "common.h"
#include "linattn_kv_common.h"
#include "tensormap.h"
#include "tile_common.cuh"

namespace dac {

constexpr uint32_t kKvSlab = kTileM * 128;     // 128 rows x 64 bf16 (16 KB)
constexpr uint32_t kColC = 256, kColS = 384;
constexpr int kCtxRec = 32 * 32 + 64;          // {C[32][32], m[32], S[32]} per (image, head)

struct KvParams {
  int tiles, tiles_per_image, stages, nbuf;
  float shift_max[4];      // per head: max_d c_d * log2(e) (from the caller's [128] bounds, read once at plan creation)
  float* ctx_acc;          // [B][4][slots][kCtxRec]: one partial record per (CTA, image), slot = CTA - first CTA of the image
  int slots;
  const float* ln_stats;   // folded PreNorm: per-pixel {mean, rstd} of the RAW input row (NULL: the input is normalised)
  const float* ln_colsum;  // [256] sum_c W'[n][c] of the bf16 weight rows, in the packed (grouped) row order
};

template <int C>
__global__ void __launch_bounds__(kThreads, 1)
linattn_kv_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW,
                  const __grid_constant__ KvParams p) {
  constexpr int kCh = C / 64;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wres = smem;                                         // [kCh][2 groups] x [128 rows x 128 B]
  uint8_t* ones = wres + kCh * 2 * kKvSlab;                     // 128 x 64 bf16 of 1.0
  uint8_t* pv = ones + kKvSlab;                                 // [nbuf][P0 P1 V0 V1]
  uint8_t* ring = pv + static_cast<size_t>(p.nbuf) * 4 * kKvSlab;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + static_cast<size_t>(p.stages) * kKvSlab);
  uint64_t* full = bars;                 // [8]
  uint64_t* empty = bars + 8;            // [8]
  uint64_t* acc_full = bars + 16;        // [2]
  uint64_t* acc_empty = bars + 18;       // [2] count 128
  uint64_t* pv_full = bars + 20;         // [2] count 256
  uint64_t* pv_free = bars + 22;         // [2]
  uint64_t* ctx_done = bars + 24;
  uint64_t* ctx_flushed = bars + 25;     // count 128
  uint64_t* w_full = bars + 26;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 27);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int begin, end;
  tile_range(p.tiles, begin, end);
  const int n = end - begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapX);
    tma_prefetch_desc(&mapW);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&acc_full[g], 1);
      mbar_init(&acc_empty[g], 128);
      mbar_init(&pv_full[g], 256);
      mbar_init(&pv_free[g], 1);
    }
    mbar_init(ctx_done, 1);
    mbar_init(ctx_flushed, 128);
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  // the tile of ones (B operand of S += P^T 1): plain stores, every element equal, so the swizzle does not matter
  for (uint32_t i = threadIdx.x; i < kKvSlab / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(ones)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch();   // the next kernel may be scheduled (this one sits behind a memset node and is launched plainly)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, kCh * 2 * kKvSlab);
      for (int ck = 0; ck < kCh; ++ck)
        for (int g = 0; g < 2; ++g)
          tma_load_2d(wres + (ck * 2 + g) * kKvSlab, &mapW, w_full, ck * 64, g * 128);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < n; ++i) {
        for (int ck = 0; ck < kCh; ++ck) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], kKvSlab);
          tma_load_2d(ring + static_cast<size_t>(stage) * kKvSlab, &mapX, &full[stage], ck * 64, (begin + i) * kTileM);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: kv(0); then per tile i: kv(i+1), ctx(i) =====================
    const uint32_t idesc_kv = make_idesc_bf16(kTileM, 128);
    const uint32_t idesc_c = make_idesc_bf16(kTileM, 128) | (1u << 15) | (1u << 16);   // A and B MN-major
    const uint32_t idesc_s = make_idesc_bf16(kTileM, 16) | (1u << 15) | (1u << 16);
    const uint64_t desc_k = make_sw128_desc(0);
    // MN-major operands spanning two 64-channel swizzle atoms: leading byte offset = distance between the atoms (slabs)
    const uint64_t desc_mn = (desc_k & ~(static_cast<uint64_t>(0x3FFF) << 16)) | (static_cast<uint64_t>(kKvSlab >> 4) << 16);
    const uint32_t ring_lo = (smem_u32(ring) & 0x3FFFF) >> 4, w_lo = (smem_u32(wres) & 0x3FFFF) >> 4,
                   pv_lo = (smem_u32(pv) & 0x3FFFF) >> 4, ones_lo = (smem_u32(ones) & 0x3FFFF) >> 4,
                   slab_lo = kKvSlab >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int cur_img = -1;
    uint32_t flushes = 0, ctx_acc_flag = 0;
    auto kv = [&](int i) {
      for (int ck = 0; ck < kCh; ++ck) {
        mbar_wait(&full[stage], phase);
        if (ck == 0) {
          mbar_wait(&acc_empty[0], (i & 1) ^ 1);
          mbar_wait(&acc_empty[1], (i & 1) ^ 1);
        }
        tc_fence_after();
        const uint64_t adesc = desc_k | (ring_lo + stage * slab_lo);
        if (elect_one()) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const uint64_t bdesc = desc_k | (w_lo + (ck * 2 + g) * slab_lo);
            const uint32_t d = tmem_base + g * 128;
            umma_bf16(d, adesc, bdesc, idesc_kv, ck ? 1u : 0u);
            umma_bf16(d, adesc + 2, bdesc + 2, idesc_kv, 1u);
            umma_bf16(d, adesc + 4, bdesc + 4, idesc_kv, 1u);
            umma_bf16(d, adesc + 6, bdesc + 6, idesc_kv, 1u);
          }
          umma_commit(&empty[stage]);
          if (ck == kCh - 1) {
            umma_commit(&acc_full[0]);
            umma_commit(&acc_full[1]);
          }
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    };
    auto ctx = [&](int i) {
      const int img = (begin + i) / p.tiles_per_image;
      if (img != cur_img) {
        if (cur_img >= 0) {               // hand the finished image's accumulator to epilogue group 0, wait for the read
          if (elect_one()) umma_commit(ctx_done);
          __syncwarp();
          mbar_wait(ctx_flushed, flushes & 1);
          tc_fence_after();
          ++flushes;
        }
        cur_img = img;
        ctx_acc_flag = 0;
      }
      const int b = p.nbuf == 2 ? (i & 1) : 0;
      const uint32_t use = p.nbuf == 2 ? (i >> 1) : i;     // how many times this buffer has been filled before
      mbar_wait(&pv_full[b], use & 1);
      tc_fence_after();
      const uint32_t p_lo = pv_lo + b * 4 * slab_lo, v_lo = p_lo + 2 * slab_lo;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {   // 16 pixels (rows of the [pixel][channel] tiles) per K step: +2048 B
          const uint64_t adesc = desc_mn | (p_lo + ks * 128);
          umma_bf16(tmem_base + kColC, adesc, desc_mn | (v_lo + ks * 128), idesc_c, ctx_acc_flag | (ks ? 1u : 0u));
          umma_bf16(tmem_base + kColS, adesc, desc_mn | (ones_lo + ks * 128), idesc_s, ctx_acc_flag | (ks ? 1u : 0u));
        }
        umma_commit(&pv_free[b]);
      }
      __syncwarp();
      ctx_acc_flag = 1;
    };
    mbar_wait(w_full, 0);
    if (n > 0) kv(0);
    for (int i = 0; i < n; ++i) {
      if (i + 1 < n) kv(i + 1);
      ctx(i);
    }
    if (n > 0) {
      if (elect_one()) umma_commit(ctx_done);
      __syncwarp();
    }
  } else {
    // ===================== epilogue groups =====================
    const int quad = warp & 3;
    const int g = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t acc = lane_base + g * 128;
    float v[32];
    uint32_t flushes = 0;
    // group 0 adds the accumulated {C, S} of image `img` into ctx_acc: TMEM lane = (head, d)
    auto flush = [&](int img) {
      mbar_wait(ctx_done, flushes & 1);
      tc_fence_after();
      ++flushes;
      const int h = quad, d = lane;                       // lane row = 32 h + d
      const int slot = static_cast<int>(blockIdx.x) - tile_owner(img * p.tiles_per_image, p.tiles, gridDim.x);
      float* rec = p.ctx_acc + ((static_cast<long long>(img) * 4 + h) * p.slots + slot) * kCtxRec;
      chunk_from_tmem(lane_base + kColC + 32 * h, v);     // the diagonal block: columns (h, e)
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(rec + d * 32 + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      uint32_t r[16];
      tmem_ld16(lane_base + kColS, r);
      tmem_ld_wait();
      rec[1056 + d] = __uint_as_float(r[0]);
      tc_fence_before();
      mbar_arrive(ctx_flushed);
    };
    int prev_img = -1;
    for (int i = 0; i < n; ++i) {
      const int img = (begin + i) / p.tiles_per_image;
      const int b = p.nbuf == 2 ? (i & 1) : 0;
      const uint32_t use = p.nbuf == 2 ? (i >> 1) : i;
      uint8_t* pt = pv + (static_cast<size_t>(b) * 4 + g) * kKvSlab;        // P slab of this group's heads
      uint8_t* vt = pt + 2 * kKvSlab;                                       // V slab
      // folded PreNorm: W' LN(x) = rstd * (W' x - mean * colsum(W')); ka / kb put that and log2(e) into the exponent's FMA
      // The shift only has to keep exp() in range: softmax over the pixels of channel d is invariant to ANY per-channel
      // constant (C and S carry the same factor), so ONE scalar - the largest bound of the head - serves all 32 channels
      // and the epilogue loads no per-channel vector at all (16 x LDG.128 per row and tile were 28 % of its stall samples).
      float ka = 1.4426950408889634f, kb = 0.f, va = 1.f, vb = 0.f;
      if (p.ln_stats) {
        const float2 ms = __ldg(reinterpret_cast<const float2*>(p.ln_stats) + static_cast<long long>(begin + i) * kTileM + row);
        va = ms.y;
        vb = -ms.x * ms.y;
        ka = va * 1.4426950408889634f;
        kb = vb * 1.4426950408889634f;
      }
      mbar_wait(&acc_full[g], i & 1);
      mbar_wait(&pv_free[b], (use & 1) ^ 1);                                // GEMM 2 of the previous user is done
      tc_fence_after();
      // v first (convert, stage), then BOTH k chunks go to registers and the TMEM stage is handed back at once: GEMM 1 of
      // the next tile runs under the exponentials / packing / staging of k instead of after them
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        chunk_from_tmem(acc + 64 + 32 * j, v);                              // v of head 2g + j
        if (p.ln_stats) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 cs = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + g * 128 + 64 + 32 * j) + q);
            v[4 * q] = fmaf(v[4 * q], va, cs.x * vb);
            v[4 * q + 1] = fmaf(v[4 * q + 1], va, cs.y * vb);
            v[4 * q + 2] = fmaf(v[4 * q + 2], va, cs.z * vb);
            v[4 * q + 3] = fmaf(v[4 * q + 3], va, cs.w * vb);
          }
        }
        chunk_stage_bf16(vt, row, 32 * j, v);
      }
      uint32_t r[2][32];
      tmem_ld32(acc, r[0]);            // k of head 2g
      tmem_ld32(acc + 32, r[1]);       // k of head 2g + 1
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&acc_empty[g]);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
#pragma unroll
        const float shj = p.shift_max[2 * g + j];
        for (int q = 0; q < 8; ++q) {
          float4 sh = make_float4(shj, shj, shj, shj);
          if (p.ln_stats) {
            const float4 cs = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + g * 128 + 32 * j) + q);
            sh.x = fmaf(cs.x, -kb, sh.x); sh.y = fmaf(cs.y, -kb, sh.y); sh.z = fmaf(cs.z, -kb, sh.z); sh.w = fmaf(cs.w, -kb, sh.w);
          }
          v[4 * q] = ex2_approx(fmaf(__uint_as_float(r[j][4 * q]), ka, -sh.x));
          v[4 * q + 1] = ex2_approx(fmaf(__uint_as_float(r[j][4 * q + 1]), ka, -sh.y));
          v[4 * q + 2] = ex2_approx(fmaf(__uint_as_float(r[j][4 * q + 2]), ka, -sh.z));
          v[4 * q + 3] = ex2_approx(fmaf(__uint_as_float(r[j][4 * q + 3]), ka, -sh.w));
        }
        chunk_stage_bf16(pt, row, 32 * j, v);
      }
      fence_proxy_async();
      mbar_arrive(&pv_full[b]);
      if (g == 0 && prev_img >= 0 && img != prev_img) flush(prev_img);      // the issuer is waiting before ctx(i)
      prev_img = img;
    }
    if (g == 0 && n > 0) flush(prev_img);
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

struct dac_kv_plan {
  CUtensorMap mapX, mapW;
  KvParams kp;
  Kv2Params kp2;
  int C, B, grid, smem;
  int prenorm;             // the input is the RAW tensor: linattn_kv2_kernel normalises its rows in shared memory
};

static int kv_encode_2d(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t rows, uint32_t box_rows,
                        const char* what) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return set_error(-10, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {inner * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(-11, "cuTensorMapEncodeTiled(%s) failed: CUresult %d", what, (int)r);
  return 0;
}

extern "C" int dac_linattn_kv_create(const void* xn, const void* wkv, const float* kv_shift, float* ctx_acc,
                                     int32_t ctx_slots, const float* ln_stats, const float* ln_colsum, int32_t B,
                                     int32_t hw, int32_t C, int32_t prenorm, float prenorm_eps, dac_kv_t* plan) {
  if (!xn || !wkv || !kv_shift || !ctx_acc || !plan) return set_error(-1, "dac_linattn_kv_create: null argument");
  *plan = nullptr;
  if (C != 64 && C != 128) return set_error(-2, "dac_linattn_kv_create: C must be 64 or 128 (got %d)", C);
  if (prenorm && (C != 64 || ln_stats))
    return set_error(-2, "dac_linattn_kv_create: in-kernel PreNorm needs C = 64 and no ln_stats");
  if (B <= 0 || hw <= 0 || hw % kTileM) return set_error(-2, "dac_linattn_kv_create: hw must be a multiple of 128");
  if ((reinterpret_cast<uintptr_t>(xn) | reinterpret_cast<uintptr_t>(wkv) | reinterpret_cast<uintptr_t>(kv_shift) |
       reinterpret_cast<uintptr_t>(ctx_acc)) & 15)
    return set_error(-2, "dac_linattn_kv_create: pointers must be 16-byte aligned");
  dac_kv_plan* pl = new (std::nothrow) dac_kv_plan();
  if (!pl) return set_error(-3, "out of host memory");
  const uint64_t rows = static_cast<uint64_t>(B) * hw;
  int rc = kv_encode_2d(&pl->mapX, xn, C, rows, kTileM, "xn");
  if (!rc) rc = kv_encode_2d(&pl->mapW, wkv, C, prenorm ? 128 : 256, 128, "wkv");
  if (rc) { delete pl; return rc; }
  KvParams& k = pl->kp;
  k.tiles = static_cast<int>(rows / kTileM);
  k.tiles_per_image = hw / kTileM;
  {
    float sh[128];
    cudaError_t ce = cudaMemcpy(sh, kv_shift, sizeof(sh), cudaMemcpyDeviceToHost);
    if (ce != cudaSuccess) {
      delete pl;
      return set_error(-20, "dac_linattn_kv_create: reading kv_shift: %s", cudaGetErrorString(ce));
    }
    for (int h = 0; h < 4; ++h) {
      float m = sh[h * 32];
      for (int d = 1; d < 32; ++d) m = sh[h * 32 + d] > m ? sh[h * 32 + d] : m;
      k.shift_max[h] = m;
    }
  }
  k.ctx_acc = ctx_acc;
  k.ln_stats = ln_stats;
  k.ln_colsum = ln_colsum;
  if ((ln_stats != nullptr) != (ln_colsum != nullptr) ||
      ((reinterpret_cast<uintptr_t>(ln_stats) | reinterpret_cast<uintptr_t>(ln_colsum)) & 15)) {
    delete pl;
    return set_error(-2, "dac_linattn_kv_create: ln_stats and ln_colsum come together, 16-byte aligned");
  }
  k.nbuf = C == 64 ? 2 : 1;
  const int fixed = (C / 64) * 2 * (int)kKvSlab + (int)kKvSlab + k.nbuf * 4 * (int)kKvSlab + 1024 + 512;
  int stages = (227 * 1024 - fixed) / (int)kKvSlab;
  if (stages > 8) stages = 8;
  if (stages < 2) { delete pl; return set_error(-2, "dac_linattn_kv_create: does not fit shared memory"); }
  k.stages = stages;
  pl->smem = fixed + stages * (int)kKvSlab;
  pl->C = C;
  pl->B = B;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  pl->grid = k.tiles < sms ? k.tiles : sms;
  k.slots = ctx_slots;
  pl->prenorm = prenorm ? 1 : 0;
  if (prenorm) {
    Kv2Params& k2 = pl->kp2;
    k2.tiles = k.tiles;
    k2.tiles_per_image = k.tiles_per_image;
    for (int h = 0; h < 4; ++h) k2.shift_max[h] = k.shift_max[h];
    k2.ctx_acc = ctx_acc;
    k2.slots = ctx_slots;
    k2.ln_eps = prenorm_eps;
    const char* dbg = getenv("DAC_KV2_DBG");
    k2.dbg = dbg ? atoi(dbg) : 0;
    k2.prof = nullptr;
  }
  if (ctx_slots < max_image_span(B, k.tiles_per_image, pl->grid)) {
    const int need = max_image_span(B, k.tiles_per_image, pl->grid);
    delete pl;
    return set_error(-2, "dac_linattn_kv_create: ctx_slots must be >= %d (dac_linattn_ctx_slots), got %d", need, ctx_slots);
  }
  cudaError_t e = C == 64 ? cudaFuncSetAttribute(linattn_kv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, pl->smem)
                          : cudaFuncSetAttribute(linattn_kv_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, pl->smem);
  if (e != cudaSuccess) {
    const int smem = pl->smem;
    delete pl;
    return set_error(-12, "dac_linattn_kv_create: cudaFuncSetAttribute(%d B smem): %s", smem, cudaGetErrorString(e));
  }
  *plan = pl;
  return 0;
}

extern "C" int dac_linattn_kv_launch(dac_kv_t pl, dac_stream_t stream) {
  if (!pl) return set_error(-1, "dac_linattn_kv_launch: null plan");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // in-kernel-PreNorm mode: every record dac_linattn_fold_g reads is written by this launch - nothing to clear
  if (pl->prenorm) return dac_kv2_launch(pl->mapX, pl->mapW, pl->kp2, pl->grid, st);
  cudaError_t e = cudaMemsetAsync(pl->kp.ctx_acc, 0, sizeof(float) * pl->B * 4 * pl->kp.slots * kCtxRec, st);
  if (e != cudaSuccess) return set_error(-20, "dac_linattn_kv_launch: memset failed: %s", cudaGetErrorString(e));
  if (pl->C == 64) linattn_kv_kernel<64><<<pl->grid, kThreads, pl->smem, st>>>(pl->mapX, pl->mapW, pl->kp);
  else linattn_kv_kernel<128><<<pl->grid, kThreads, pl->smem, st>>>(pl->mapX, pl->mapW, pl->kp);
  return check_launch("linattn_kv_kernel");
}

extern "C" void dac_linattn_kv_destroy(dac_kv_t pl) { delete pl; }
