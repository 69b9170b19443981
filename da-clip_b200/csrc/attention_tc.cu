// Softmax self-attention with head dim 32 on the tcgen05 tensor cores (BasicTransformerBlock.attn1 of the
// SpatialTransformer, attention.py:152-193 of the reference: 1024 tokens at 256^2 images, 4096 at 512^2).
//
// One work item = (image, PAIR of heads, 128 query rows).  q | k | v are packed along channels, 32 per head, so one
// 128B-swizzled TMA box of 64 channels carries two heads; head A contracts over K-steps 0-1 of the box, head B over
// K-steps 2-3 (descriptor start + 64 B).  Per 128-key block and head:
//   S = Q K^T            tcgen05.mma  M128 x N128 x K32  -> TMEM (fp32)
//   P = exp2((S - m) c)  the head's softmax group (4 warps, thread = query row: row max / sum need no shuffles),
//                        bf16 into shared memory in the K-major 128B-swizzled A-operand layout
//   O_blk = P V          tcgen05.mma  M128 x N64 x K128, V as an MN-major B operand straight from its [key][channel]
//                        box (N = both heads' 64 channels so the operand is a whole swizzle atom; a head keeps its 32)
//   O = O * alpha + O_blk   in the group's registers (32 fp32 per thread), from a 32-column tcgen05.ld
// The kernel is bound by the 128 exponentials per (query, key block) on the SFU pipe (16/clk/SM); the tensor pipe and
// TMA run underneath: while one group is in its exponentials the issuer computes the other head's S and P V.
// Roles: warp 0 = TMA producer, warps 1 and 10 = MMA issuers of head A / B, warps 2-5 = softmax of head A, 6-9 = head B.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdlib.h>

#include "../../include/dac_b200.h"
#include "common.h"
#include "tensormap.h"
#include "tile_common.cuh"

namespace dac {

constexpr uint32_t kAtSlab = kTileM * 128;   // 128 rows x 64 bf16 (16 KB)
constexpr int kAtStages = 4;                 // K / V ring (two key blocks in flight)
constexpr uint32_t kAtColS = 0;              // TMEM: S_A [0,128), S_B [128,256)
constexpr uint32_t kAtColO = 256;            //       O_blk A [256,320), O_blk B [320,384)

struct AttnParams {
  int items, q_tiles, pairs, n, heads;
  float scale_log2;                          // d^-0.5 * log2(e)
  __nv_bfloat16* out;
  int dbg;                                   // profiling only (DAC_ATTN_DEBUG): 1 no exponentials, 2 no max pass, 4 no P store
};

constexpr int kAtThreads = 352;   // warp 0 TMA, warps 1 / 10 MMA issuers (one per head), warps 2-9 softmax

__global__ void __launch_bounds__(kAtThreads, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap mapQKV, const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* qs = smem;                                  // Q tile of the current item
  uint8_t* ring = qs + kAtSlab;                        // [kAtStages] K / V tiles, alternating
  uint8_t* ps = ring + kAtStages * kAtSlab;            // [2 heads][2 buffers][2 slabs] P tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(ps + 8 * kAtSlab);
  uint64_t* full = bars;                               // [kAtStages]
  uint64_t* empty = bars + 8;                          // [kAtStages]
  uint64_t* q_full = bars + 16;
  uint64_t* q_free = bars + 17;
  uint64_t* s_full = bars + 18;                        // [2]  S of the head computed
  uint64_t* s_free = bars + 20;                        // [2]  ... read by its group (count 128)
  uint64_t* p_full = bars + 22;                        // [2]  P tile staged (count 128)
  uint64_t* o_full = bars + 24;                        // [2]  O_blk computed (also: P tile consumed)
  uint64_t* o_free = bars + 26;                        // [2]  ... read by its group (count 128)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int begin, end;
  tile_range(p.items, begin, end);
  const int kblocks = p.n / kTileM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapQKV);
    for (int s = 0; s < kAtStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 2);      // released by both heads' issuers
    }
    mbar_init(q_full, 1);
    mbar_init(q_free, 2);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&s_free[g], 128);
      mbar_init(&p_full[g], 128);
      mbar_init(&o_full[g], 1);
      mbar_init(&o_free[g], 128);
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
  griddep_launch();   // programmatic dependent launch: see conv_kernel.cuh

  // item -> (image b, head pair, query tile); rows of the packed [B*n, 3*heads*32] matrix
  auto decode = [&](int item, int& row0, int& pair, int& brow) {
    const int qt = item % p.q_tiles;
    const int r = item / p.q_tiles;
    pair = r % p.pairs;
    const int b = r / p.pairs;
    brow = b * p.n;
    row0 = brow + qt * kTileM;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      griddep_wait();
      for (int it = begin; it < end; ++it) {
        int row0, pair, brow;
        decode(it, row0, pair, brow);
        const int li = it - begin;
        mbar_wait(q_free, (li & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, kAtSlab);
        tma_load_2d(qs, &mapQKV, q_full, pair * 64, row0);
        for (int j = 0; j < kblocks; ++j) {
          for (int kv = 0; kv < 2; ++kv) {          // K tile, then V tile of key block j
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], kAtSlab);
            tma_load_2d(ring + stage * kAtSlab, &mapQKV, &full[stage], (1 + kv) * p.heads * 32 + pair * 64,
                        brow + j * kTileM);
            if (++stage == kAtStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 10) {
    // ===================== MMA issuers: one warp PER HEAD (warp 1 = head A, warp 10 = head B) =====================
    // Each walks the K / V ring on its own and serves only its softmax group, so the two heads drift out of phase
    // instead of marching in lockstep behind one in-order issuer (one group's exponentials then overlap the other
    // group's TMEM / barrier latencies).  Ring stages and the Q tile are released by both (barrier count 2).
    const int g = warp == 1 ? 0 : 1;
    const uint32_t idesc_s = make_idesc_bf16(kTileM, 128);
    const uint32_t idesc_o = make_idesc_bf16(kTileM, 64) | (1u << 16);      // B = V is MN-major
    const uint64_t desc_k = make_sw128_desc(0);                              // K-major, 1024 B between 8-row groups
    const uint32_t qs_lo = (smem_u32(qs) & 0x3FFFF) >> 4, ring_lo = (smem_u32(ring) & 0x3FFFF) >> 4,
                   ps_lo = (smem_u32(ps) & 0x3FFFF) >> 4, slab_lo = kAtSlab >> 4;
    const uint32_t d_s = tmem_base + kAtColS + g * 128, d_o = tmem_base + kAtColO + g * 64;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t blk = 0;                       // key blocks issued so far (phase of the per-head barriers)
    // P V of a block is issued one block late: its softmax runs while the next block's S is computed
    struct Pending { int stage; uint32_t vphase; uint32_t blk; bool valid; } pend = {0, 0, 0, false};
    auto pv = [&]() {
      mbar_wait(&full[pend.stage], pend.vphase);
      mbar_wait(&p_full[g], pend.blk & 1);
      mbar_wait(&o_free[g], (pend.blk & 1) ^ 1);
      tc_fence_after();
      const uint32_t v_lo = ring_lo + pend.stage * slab_lo;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          // A: P rows x 16 keys (K-major, slab ks / 4, 32 B per K step); B: 16 key rows of V (MN-major: +2048 B)
          const uint64_t adesc = desc_k | (ps_lo + ((g * 2 + (pend.blk & 1)) * 2 + (ks >> 2)) * slab_lo + (ks & 3) * 2);
          const uint64_t bdesc = desc_k | (v_lo + ks * 128);
          umma_bf16(d_o, adesc, bdesc, idesc_o, ks ? 1u : 0u);
        }
        umma_commit(&o_full[g]);
        umma_commit(&empty[pend.stage]);
      }
      __syncwarp();
      pend.valid = false;
    };
    for (int it = begin; it < end; ++it) {
      const int li = it - begin;
      mbar_wait(q_full, li & 1);
      tc_fence_after();
      for (int j = 0; j < kblocks; ++j, ++blk) {
        // ---- S = Q K^T of key block j for this head (K-steps 2g, 2g + 1 of the 64-channel boxes)
        mbar_wait(&full[stage], phase);
        mbar_wait(&s_free[g], (blk & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_lo = ring_lo + stage * slab_lo;
        if (elect_one()) {
          umma_bf16(d_s, desc_k | (qs_lo + g * 4), desc_k | (k_lo + g * 4), idesc_s, 0u);
          umma_bf16(d_s, desc_k | (qs_lo + g * 4 + 2), desc_k | (k_lo + g * 4 + 2), idesc_s, 1u);
          umma_commit(&s_full[g]);
          umma_commit(&empty[stage]);
          if (j == kblocks - 1) umma_commit(q_free);      // last S of the item: the Q tile may be replaced
        }
        __syncwarp();
        if (++stage == kAtStages) {
          stage = 0;
          phase ^= 1;
        }
        // ---- P V of the previous block (its softmax ran while this block's S was computed)
        if (pend.valid) pv();
        pend.stage = stage;
        pend.vphase = phase;
        pend.blk = blk;
        pend.valid = true;
        if (++stage == kAtStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    if (pend.valid) pv();
  } else {
    // ===================== softmax groups: group g = head g of the pair, thread = query row =====================
    const int quad = warp & 3;
    const int g = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    griddep_wait();
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t s_addr = lane_base + kAtColS + g * 128;
    const uint32_t o_addr = lane_base + kAtColO + g * 64 + g * 32;   // this head's 32 channels of the N = 64 product
    const float c = p.scale_log2;
    uint32_t blk = 0;
    float v[32];
    for (int it = begin; it < end; ++it) {
      int row0, pair, brow;
      decode(it, row0, pair, brow);
      float o[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = 0.f;
      float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;
      for (int j = 0; j < kblocks; ++j, ++blk) {
        mbar_wait(&s_full[g], blk & 1);
        tc_fence_after();
        // pass A: row max of this key block (two TMEM loads in flight, four independent max chains)
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        if (p.dbg & 2) m4[0] = 0.f;
        else
#pragma unroll
        for (int cc = 0; cc < 128; cc += 64) {
          uint32_t r0[32], r1[32];
          tmem_ld32(s_addr + cc, r0);
          tmem_ld32(s_addr + cc + 32, r1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(r0[i]));
            m4[(i + 2) & 3] = fmaxf(m4[(i + 2) & 3], __uint_as_float(r1[i]));
          }
        }
        const float m_blk = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        const float m_new = fmaxf(m_run, m_blk);
        const float alpha = ex2_approx((m_run - m_new) * c);      // 0 on the first block (m_run = -inf)
        const float mc = m_new * c;
        // pass B: P = exp2(S c - m c), row sum (four independent chains), bf16 A operand of P V
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
        uint8_t* pt = ps + (g * 2 + (blk & 1)) * 2 * kAtSlab;     // P tiles are double-buffered per head
#pragma unroll
        for (int cc = 0; cc < 128; cc += 32) {
          chunk_from_tmem(s_addr + cc, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v[i] = (p.dbg & 1) ? fmaf(v[i], c, -mc) : ex2_approx(fmaf(v[i], c, -mc));
            l4[i & 3] += v[i];
          }
          if (!(p.dbg & 4)) chunk_stage_bf16(pt, row, cc, v);
        }
        const float l_blk = (l4[0] + l4[1]) + (l4[2] + l4[3]);
        tc_fence_before();
        mbar_arrive(&s_free[g]);                 // S read twice, done: the issuer may overwrite it
        fence_proxy_async();                     // P (generic-proxy stores) -> visible to the tensor core
        mbar_arrive(&p_full[g]);
        l_run = fmaf(l_run, alpha, l_blk);
        m_run = m_new;
        // fold the PREVIOUS block's P V into O: its MMAs ran under this block's two passes (o_full also means that
        // block's P buffer is free again, which the block after this one relies on)
        if (j > 0) {
          mbar_wait(&o_full[g], (blk - 1) & 1);
          tc_fence_after();
          chunk_from_tmem(o_addr, v);
          tc_fence_before();
          mbar_arrive(&o_free[g]);
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = fmaf(o[i], alpha_prev, v[i]);
        }
        alpha_prev = alpha;
      }
      // last block's P V, then normalise and store this head's 32 channels of the 128 query rows
      mbar_wait(&o_full[g], (blk - 1) & 1);
      tc_fence_after();
      chunk_from_tmem(o_addr, v);
      tc_fence_before();
      mbar_arrive(&o_free[g]);
      const float inv = 1.0f / l_run;
      // careful: o holds sum over blocks < last scaled up to m of block last-1; alpha_prev rescales to the final max
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaf(o[i], alpha_prev, v[i]) * inv;
      __nv_bfloat16* dst = p.out + (static_cast<int64_t>(row0) + row) * (p.heads * 32) + (pair * 2 + g) * 32;
      chunk_store_bf16(dst, v);
    }
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

// Called by dac_attention (attention.cu) for d = 32, n % 128 == 0, even head count.
int dac_attention_tc(const void* qkv, void* out, int B, int n, int heads, cudaStream_t stream) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return set_error(-10, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 31)
    return set_error(-2, "dac_attention: pointers must be 32-byte aligned");
  CUtensorMap map;
  const uint64_t cols = 3ull * heads * 32;
  cuuint64_t dims[2] = {cols, static_cast<cuuint64_t>(B) * n};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(kTileM)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(qkv), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(-11, "cuTensorMapEncodeTiled(qkv) failed: CUresult %d", (int)r);
  AttnParams k;
  k.q_tiles = n / kTileM;
  k.pairs = heads / 2;
  k.items = B * k.pairs * k.q_tiles;
  k.n = n;
  k.heads = heads;
  k.scale_log2 = 0.17677669529663687f * 1.4426950408889634f;
  k.out = static_cast<__nv_bfloat16*>(out);
  k.dbg = getenv("DAC_ATTN_DEBUG") ? atoi(getenv("DAC_ATTN_DEBUG")) : 0;
  const int smem = (1 + kAtStages + 8) * (int)kAtSlab + 1024 + 512;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(-12, "dac_attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = k.items < sms ? k.items : sms;
  launch_k(attn_tc_kernel, dim3(grid), dim3(kAtThreads), smem, stream, map, k);
  return check_launch("attn_tc_kernel");
}
