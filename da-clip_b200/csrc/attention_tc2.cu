// Softmax self-attention with head dim 32 on the tcgen05 tensor cores, second generation (round 2).
// (BasicTransformerBlock.attn1 of the SpatialTransformer, attention.py:152-193 of the reference: 1024 tokens at 256^2
// images, 4096 at 512^2.)
//
// What round 1's kernel (attention_tc.cu) measured: with the exponentials, the row-max pass AND the P stores removed it
// still took 85 % of its time (profiles/r02_epilogue_decomposition.txt) - it is not bound by the SFU pipe but by the
// per-block dependency chain of ONE softmax warp per (head, TMEM lane quadrant): S MMA -> barrier -> 2 TMEM sweeps of 128
// columns -> P staged -> barrier -> P V MMA -> barrier -> O read back and folded, ~300 instructions per row and block
// at two warps per scheduler.  This version
//   * splits every head's KEYS between two independent softmax streams: stream (g, h) owns keys [64 h, 64 h + 64) of each
//     128-key block, its own running (max, sum) per row and its own O accumulator; four streams = 16 softmax warps, four
//     per scheduler, and no exchange between the halves until the end of the work item (one merge per 1024 / 4096 keys);
//   * reads S once (64 columns per thread stay in registers between the max and the exponentials);
//   * leaves O in TENSOR MEMORY: P V accumulates across key blocks on the tensor core, and a row is rescaled in place
//     (tcgen05.ld -> multiply -> tcgen05.st) only when its running max grew by more than 2^8 (the reference max is kept
//     otherwise, so P may reach 2^8 instead of 1: harmless in bf16 / fp32) - no per-block read-back, no fold.
// One work item = (image, PAIR of heads, 128 query rows).  q | k | v are packed along channels, 32 per head, so one
// 128B-swizzled TMA box of 64 channels carries two heads.  Per 128-key block and head g:
//   S_g = Q K^T                 tcgen05.mma M128 x N128 x K32 -> TMEM columns [128 g, 128 g + 128)
//   P_gh = exp2((S - m) c)      stream (g, h): bf16, K-major 128B-swizzled A-operand tile [128 rows][64 keys] in smem
//   O_gh += P_gh V[64 h ..]     tcgen05.mma M128 x N64 x K64, V as an MN-major B operand straight from its [key][channel]
//                               box -> TMEM columns [256 + 64 (2 g + h), +64) (the head keeps its 32 of the 64 channels)
// Roles (20 warps): warp 0 TMA producer, warps 1 / 2 MMA issuers of head A / B, warp 3 idle (keeps warp % 4 = TMEM lane
// quadrant for the rest), warps 4-19 the four softmax streams.
#ifdef DAC_ATTN2_DEBUG
#define DAC_MBAR_DEBUG 1
#endif
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <stdio.h>

#include "../../include/dac_b200.h"
#include "common.h"
#include "tensormap.h"
#include "tile_common.cuh"

namespace dac {

// DAC_ATTN2_PROF (hand-built debug library only): per-warp cycle totals spent in each kind of barrier wait
#ifdef DAC_ATTN2_PROF
#define A2_WAIT(bar, parity, slot)                                   \
  do {                                                               \
    const long long t0_ = clock64();                                 \
    mbar_wait(bar, parity);                                          \
    prof_acc[slot] += clock64() - t0_;                               \
  } while (0)
#else
#define A2_WAIT(bar, parity, slot) mbar_wait(bar, parity)
#endif

constexpr uint32_t kA2Slab = kTileM * 128;   // 128 rows x 64 bf16 (16 KB)
constexpr int kA2Stages = 4;                 // K / V ring (two key blocks in flight)
constexpr uint32_t kA2ColS = 0;              // TMEM: S_A [0,128), S_B [128,256)
constexpr uint32_t kA2ColO = 256;            //       O of stream (g, h) at 256 + 64 (2 g + h)
constexpr int kA2Threads = 640;
constexpr float kA2Lazy = 8.0f;              // rescale O only when the row max (in log2 units) grew by more than this

struct Attn2Params {
  int items, q_tiles, pairs, n, heads;
  float scale_log2;                          // d^-0.5 * log2(e)
  __nv_bfloat16* out;
  long long* prof;                           // DAC_ATTN2_PROF: [20 warps][8] cycles (CTA 0)
};

// exp2 of a pair of arguments on the FMA / ALU pipes (no MUFU): round to nearest integer with the 1.5 * 2^23 trick, degree-3
// minimax polynomial of 2^f on [-0.5, 0.5] (relative error 7.5e-5, far below the bf16 rounding of P), exponent added to the
// bit pattern.  Arguments below -126 are clamped (the result would wrap; 2^-126 rounds to zero weight anyway).
__device__ __forceinline__ void ex2_poly2(uint64_t x2, float& e0, float& e1) {
  float a0, a1;
  unpack_f32x2(x2, a0, a1);
  x2 = pack_f32x2(fmaxf(a0, -126.f), fmaxf(a1, -126.f));
  const uint64_t r2 = add_f32x2(x2, pack_f32x2(12582912.f, 12582912.f));
  const uint64_t j2 = add_f32x2(r2, pack_f32x2(-12582912.f, -12582912.f));
  const uint64_t f2 = fma_f32x2(j2, pack_f32x2(-1.f, -1.f), x2);
  uint64_t p2 = fma_f32x2(pack_f32x2(0.0551716648f, 0.0551716648f), f2, pack_f32x2(0.2426111251f, 0.2426111251f));
  p2 = fma_f32x2(p2, f2, pack_f32x2(0.6932609677f, 0.6932609677f));
  p2 = fma_f32x2(p2, f2, pack_f32x2(0.9999280572f, 0.9999280572f));
  float p0, p1, r0, r1;
  unpack_f32x2(p2, p0, p1);
  unpack_f32x2(r2, r0, r1);
  e0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(r0) << 23));
  e1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(r1) << 23));
}

// which pairs of a 32-element chunk take the polynomial: 1 = every 8th, 2 = every 4th, 3 = two adjacent of 8, 4 = every 2nd
template <int POLY>
__device__ __forceinline__ constexpr bool poly_pair(int pr) {
  return POLY == 1 ? (pr & 7) == 0 : POLY == 2 ? (pr & 3) == 0 : POLY == 3 ? (pr & 7) < 2 : POLY == 4 ? (pr & 1) == 0 : false;
}

// POLY: of every 8 pairs of exponentials, this many are computed by ex2_poly2 instead of MUFU.EX2 (the kernel is bound by the
// MUFU pipe, 16 results / clk / SM, while half the issue slots idle)
template <int POLY>
__global__ void __launch_bounds__(kA2Threads, 1)
attn_tc2_kernel(const __grid_constant__ CUtensorMap mapQKV, const __grid_constant__ Attn2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* qs = smem;                                  // Q tile of the current item
  uint8_t* ring = qs + kA2Slab;                        // [kA2Stages] K / V tiles, alternating
  uint8_t* ps = ring + kA2Stages * kA2Slab;            // [4 streams][2 buffers] P tiles (one slab each)
  uint64_t* bars = reinterpret_cast<uint64_t*>(ps + 8 * kA2Slab);
  uint64_t* full = bars;                               // [kA2Stages]
  uint64_t* empty = bars + 4;                          // [kA2Stages]
  uint64_t* q_full = bars + 8;
  uint64_t* q_free = bars + 9;
  uint64_t* s_full = bars + 10;                        // [2]  S of the head computed
  uint64_t* s_free = bars + 12;                        // [2]  ... read by both of its streams (count 256)
  uint64_t* p_full = bars + 14;                        // [4]  P tile of the stream staged (count 128)
  uint64_t* o_full = bars + 18;                        // [4]  P V of the stream's block done (also: P tile consumed)
  uint64_t* o_free = bars + 22;                        // [4]  final O of the item read by the stream (count 128)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int begin, end;
  tile_range(p.items, begin, end);
  const int kblocks = p.n / kTileM;
#ifdef DAC_ATTN2_PROF
  long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long prof_t0 = clock64();
#endif

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapQKV);
    for (int s = 0; s < kA2Stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 2);      // released by both heads' issuers
    }
    mbar_init(q_full, 1);
    mbar_init(q_free, 2);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&s_free[g], 256);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&p_full[s], 128);
      mbar_init(&o_full[s], 1);
      mbar_init(&o_free[s], 128);
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
        A2_WAIT(q_free, (li & 1) ^ 1, 0);
        mbar_arrive_expect_tx(q_full, kA2Slab);
        tma_load_2d(qs, &mapQKV, q_full, pair * 64, row0);
        for (int j = 0; j < kblocks; ++j) {
          for (int kv = 0; kv < 2; ++kv) {          // K tile, then V tile of key block j
            A2_WAIT(&empty[stage], phase ^ 1, 1);
            mbar_arrive_expect_tx(&full[stage], kA2Slab);
            tma_load_2d(ring + stage * kA2Slab, &mapQKV, &full[stage], (1 + kv) * p.heads * 32 + pair * 64,
                        brow + j * kTileM);
            if (++stage == kA2Stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ===================== MMA issuers: one warp PER HEAD =====================
    // Each walks the K / V ring on its own and serves only its two softmax streams, so the heads drift out of phase
    // instead of marching in lockstep behind one in-order issuer.  Ring stages and the Q tile are released by both
    // (barrier count 2).
    const int g = warp - 1;
    const uint32_t idesc_s = make_idesc_bf16(kTileM, 128);
    const uint32_t idesc_o = make_idesc_bf16(kTileM, 64) | (1u << 16);      // B = V is MN-major
    const uint64_t desc_k = make_sw128_desc(0);                              // K-major, 1024 B between 8-row groups
    const uint32_t qs_lo = (smem_u32(qs) & 0x3FFFF) >> 4, ring_lo = (smem_u32(ring) & 0x3FFFF) >> 4,
                   ps_lo = (smem_u32(ps) & 0x3FFFF) >> 4, slab_lo = kA2Slab >> 4;
    const uint32_t d_s = tmem_base + kA2ColS + g * 128;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t blk = 0;                       // key blocks issued so far (phase of the per-stream barriers)
    // P V of a block is issued one block late: its softmax runs while the next block's S is computed
    struct Pending { int stage; uint32_t vphase; uint32_t blk; uint32_t acc; uint32_t li; bool valid; } pend =
        {0, 0, 0, 0, 0, false};
    auto pv = [&]() {
      A2_WAIT(&full[pend.stage], pend.vphase, 0);
      const uint32_t v_lo = ring_lo + pend.stage * slab_lo;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int st = g * 2 + h;
        A2_WAIT(&p_full[st], pend.blk & 1, 1 + h);
        // first block of an item overwrites O: the stream must have read the previous item's result
        if (!pend.acc) A2_WAIT(&o_free[st], (pend.li & 1) ^ 1, 3);
        tc_fence_after();
        const uint32_t d_o = tmem_base + kA2ColO + st * 64;
        const uint32_t p_lo = ps_lo + (st * 2 + (pend.blk & 1)) * slab_lo;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            // A: P rows x 16 keys (K-major, 32 B per K step); B: 16 key rows of V (MN-major: 2048 B per K step)
            const uint64_t adesc = desc_k | (p_lo + ks * 2);
            const uint64_t bdesc = desc_k | (v_lo + (h * 4 + ks) * 128);
            umma_bf16(d_o, adesc, bdesc, idesc_o, (pend.acc || ks) ? 1u : 0u);
          }
          umma_commit(&o_full[st]);
          if (h == 1) umma_commit(&empty[pend.stage]);
        }
        __syncwarp();
      }
      pend.valid = false;
    };
    for (int it = begin; it < end; ++it) {
      const int li = it - begin;
      A2_WAIT(q_full, li & 1, 4);
      tc_fence_after();
      for (int j = 0; j < kblocks; ++j, ++blk) {
        // ---- S = Q K^T of key block j for this head (K-steps 2g, 2g + 1 of the 64-channel boxes)
        A2_WAIT(&full[stage], phase, 5);
        A2_WAIT(&s_free[g], (blk & 1) ^ 1, 6);
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
        if (++stage == kA2Stages) {
          stage = 0;
          phase ^= 1;
        }
        // ---- P V of the previous block (its softmax ran while this block's S was computed)
        if (pend.valid) pv();
        pend.stage = stage;
        pend.vphase = phase;
        pend.blk = blk;
        pend.acc = j ? 1u : 0u;
        pend.li = static_cast<uint32_t>(li);
        pend.valid = true;
        if (++stage == kA2Stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    if (pend.valid) pv();
  } else if (warp >= 4) {
    // ===================== softmax streams: stream = (head g, key half h), thread = query row =====================
    const int quad = warp & 3;
    const int st = (warp - 4) >> 2;          // 0..3
    const int g = st >> 1, h = st & 1;
    const int row = quad * 32 + lane;
    griddep_wait();
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t s_addr = lane_base + kA2ColS + g * 128 + h * 64;
    const uint32_t o_addr = lane_base + kA2ColO + st * 64 + g * 32;   // this head's 32 channels of the N = 64 product
    const float c = p.scale_log2;
    const uint32_t ps_s = smem_u32(ps);
    uint32_t blk = 0;
    // scratch of the end-of-item merge: stream (g, 1) parks {O[32], m, l} of its rows in its own (idle) P buffers
    float* merge_o = reinterpret_cast<float*>(ps + ((g * 2 + 1) * 2 + 0) * kA2Slab);     // [32 ch][128 rows]
    float* merge_ml = reinterpret_cast<float*>(ps + ((g * 2 + 1) * 2 + 1) * kA2Slab);    // [2][128 rows]
    for (int it = begin; it < end; ++it) {
      int row0, pair, brow;
      decode(it, row0, pair, brow);
      float m_ref = -INFINITY, l_run = 0.f;
      for (int j = 0; j < kblocks; ++j, ++blk) {
        A2_WAIT(&s_full[g], blk & 1, 0);
        tc_fence_after();
#ifdef DAC_ATTN2_PROF
        const long long tp0 = clock64();
#endif
        float s0[32], s1[32];
        {
          uint32_t r0[32], r1[32];
          tmem_ld32(s_addr, r0);
          tmem_ld32(s_addr + 32, r1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            s0[i] = __uint_as_float(r0[i]);
            s1[i] = __uint_as_float(r1[i]);
          }
        }
        tc_fence_before();
        mbar_arrive(&s_free[g]);                 // S is in registers: the issuer may overwrite it
#ifdef DAC_ATTN2_PROF
        const long long tp1 = clock64();
        prof_acc[4] += tp1 - tp0;
#endif
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          m4[i & 3] = fmaxf(m4[i & 3], s0[i]);
          m4[(i + 2) & 3] = fmaxf(m4[(i + 2) & 3], s1[i]);
        }
        const float m_blk = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        // lazy rescale: keep the reference max unless this block exceeds it by more than 2^kA2Lazy
        const bool grow = (m_blk - m_ref) * c > kA2Lazy;       // true on the first block (m_ref = -inf)
        const float m_new = grow ? m_blk : m_ref;
        if (j > 0 && __any_sync(0xffffffffu, grow)) {
          // O (accumulated by the tensor core up to block j - 1) is rescaled in tensor memory; rows that keep their
          // reference multiply by 1.  P V of block j - 1 must have completed; P V of block j waits for p_full below.
          const float alpha = grow ? ex2_approx((m_ref - m_new) * c) : 1.0f;
          A2_WAIT(&o_full[st], (blk - 1) & 1, 1);
          tc_fence_after();
#pragma unroll
          for (int half = 0; half < 2; ++half) {   // 16 columns at a time: the 64 scores of this block stay in registers
            uint32_t o[16];
            tmem_ld16(o_addr + 16 * half, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(o_addr + 16 * half, o);
          }
          tmem_st_wait();
          tc_fence_before();
          l_run *= alpha;
        }
        m_ref = m_new;
        const float mc = m_new * c;
#ifdef DAC_ATTN2_PROF
        const long long tp2 = clock64();
        prof_acc[5] += tp2 - tp1;
#endif
        // the P buffer of this block was last read by P V of block j - 2: done before block j - 1's wait below returned
        const uint32_t pt = ps_s + (st * 2 + (blk & 1)) * kA2Slab;
        // exponent arguments and row sums as packed fp32 pairs (FFMA2 / FADD2): half the issue slots of the scalar forms
        const uint64_t c2 = pack_f32x2(c, c), mc2 = pack_f32x2(-mc, -mc);
        uint64_t lsum[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const uint64_t x2 = fma_f32x2(pack_f32x2(s0[i], s0[i + 1]), c2, mc2);
          if (poly_pair<POLY>(i >> 1)) {
            ex2_poly2(x2, s0[i], s0[i + 1]);
          } else {
            float a0, a1;
            unpack_f32x2(x2, a0, a1);
            s0[i] = ex2_approx(a0);
            s0[i + 1] = ex2_approx(a1);
          }
          lsum[(i >> 1) & 3] = add_f32x2(lsum[(i >> 1) & 3], pack_f32x2(s0[i], s0[i + 1]));
        }
        chunk_stage_bf16_s(pt, row, 0, s0);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const uint64_t x2 = fma_f32x2(pack_f32x2(s1[i], s1[i + 1]), c2, mc2);
          if (poly_pair<POLY>(i >> 1)) {
            ex2_poly2(x2, s1[i], s1[i + 1]);
          } else {
            float a0, a1;
            unpack_f32x2(x2, a0, a1);
            s1[i] = ex2_approx(a0);
            s1[i + 1] = ex2_approx(a1);
          }
          lsum[(i >> 1) & 3] = add_f32x2(lsum[(i >> 1) & 3], pack_f32x2(s1[i], s1[i + 1]));
        }
        chunk_stage_bf16_s(pt, row, 32, s1);
        float l4[4];
        {
          float x0, x1;
          unpack_f32x2(add_f32x2(add_f32x2(lsum[0], lsum[1]), add_f32x2(lsum[2], lsum[3])), x0, x1);
          l4[0] = x0; l4[1] = x1; l4[2] = 0.f; l4[3] = 0.f;
        }
        l_run += (l4[0] + l4[1]) + (l4[2] + l4[3]);
#ifdef DAC_ATTN2_PROF
        prof_acc[6] += clock64() - tp2;
#endif
        // P V of the PREVIOUS block has (long) completed by now.  The wait must come BEFORE this block's p_full arrival:
        // an mbarrier parity wait can only tell the current phase from the one before it, and once p_full(j) is complete
        // the issuer may finish P V (j) as well - a thread that tested o_full(j - 1) only then would see the parity of
        // phase j + 1 and wait for a P V that needs its own next arrival (deadlock seen at 32 key blocks).  It also
        // guarantees that block j + 1 may overwrite the P buffer block j - 1 used.
        if (j > 0) A2_WAIT(&o_full[st], (blk - 1) & 1, 2);
        fence_proxy_async();                     // P (generic-proxy stores) -> visible to the tensor core
        mbar_arrive(&p_full[st]);
      }
      // ---- end of the item: O of this stream is complete in tensor memory
      A2_WAIT(&o_full[st], (blk - 1) & 1, 3);
      tc_fence_after();
      float o[32];
      chunk_from_tmem(o_addr, o);
      tc_fence_before();
      mbar_arrive(&o_free[st]);
      // merge the two key halves of every row: the (g, 1) stream parks its state, the (g, 0) stream combines and stores
      if (h == 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i) merge_o[i * kTileM + row] = o[i];
        merge_ml[row] = m_ref;
        merge_ml[kTileM + row] = l_run;
      }
      asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");
      if (h == 0) {
        const float m1 = merge_ml[row], l1 = merge_ml[kTileM + row];
        const float m = fmaxf(m_ref, m1);
        const float w0 = ex2_approx((m_ref - m) * c), w1 = ex2_approx((m1 - m) * c);
        const float inv = 1.0f / fmaf(l_run, w0, l1 * w1);
        const float a0 = w0 * inv, a1 = w1 * inv;
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = fmaf(o[i], a0, merge_o[i * kTileM + row] * a1);
        __nv_bfloat16* dst = p.out + (static_cast<int64_t>(row0) + row) * (p.heads * 32) + (pair * 2 + g) * 32;
        chunk_store_bf16(dst, o);
      }
      // the scratch lives in the (g, 1) stream's P buffers: it may only be rewritten (next item, block 0 / 1) after the read
      asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");
    }
  }

#ifdef DAC_ATTN2_PROF
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

// Called by dac_attention (attention.cu) for d = 32, n % 128 == 0, even head count.
int dac_attention_tc2(const void* qkv, void* out, int B, int n, int heads, cudaStream_t stream) {
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
  Attn2Params k;
  k.q_tiles = n / kTileM;
  k.pairs = heads / 2;
  k.items = B * k.pairs * k.q_tiles;
  k.n = n;
  k.heads = heads;
  k.scale_log2 = 0.17677669529663687f * 1.4426950408889634f;
  k.out = static_cast<__nv_bfloat16*>(out);
  k.prof = nullptr;
#ifdef DAC_ATTN2_PROF
  static long long* prof_buf = nullptr;
  if (!prof_buf) cudaMallocManaged(&prof_buf, 20 * 8 * sizeof(long long));
  k.prof = prof_buf;
  if (getenv("DAC_ATTN2_PROF_DUMP")) {      // print the totals of the PREVIOUS launch
    cudaDeviceSynchronize();
    const char* names[3][8] = {{"q_free", "empty", "-", "-", "-", "-", "-", "total"},
                               {"full(V)", "p_full h0", "p_full h1", "o_free", "q_full", "full(K)", "s_free", "total"},
                               {"s_full", "o_full(rescale)", "o_full(loop)", "o_full(end)", "tmem_ld", "max", "exp+stage", "total"}};
    for (int w = 0; w < 20; ++w) {
      if (w == 3) continue;
      const int role = w == 0 ? 0 : (w < 3 ? 1 : 2);
      printf("warp %2d:", w);
      for (int i = 0; i < 8; ++i)
        if (names[role][i][0] != '-') printf("  %s=%lld", names[role][i], prof_buf[w * 8 + i]);
      printf("\n");
    }
  }
#endif
  const int smem = (1 + kA2Stages + 8) * (int)kA2Slab + 1024 + 512;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_tc2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(-12, "dac_attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = k.items < sms ? k.items : sms;
  if (getenv("DAC_ATTN_GRID")) grid = atoi(getenv("DAC_ATTN_GRID")) < grid ? atoi(getenv("DAC_ATTN_GRID")) : grid;   // tests: many items per CTA
  // a quarter of the exponentials on the FMA / ALU pipes: 116 -> 110 us at 16 x 16 heads x 1024 tokens, 732 -> 687 us at 4096
  // (12.5 % gives the same, 50 % is slower: the kernel then runs out of issue slots); DAC_ATTN_POLY=0 selects the all-MUFU form
  const bool poly = !getenv("DAC_ATTN_POLY") || atoi(getenv("DAC_ATTN_POLY")) != 0;
  if (poly) launch_k(attn_tc2_kernel<3>, dim3(grid), dim3(kA2Threads), smem, stream, map, k);
  else launch_k(attn_tc2_kernel<0>, dim3(grid), dim3(kA2Threads), smem, stream, map, k);
  return check_launch("attn_tc2_kernel");
}
