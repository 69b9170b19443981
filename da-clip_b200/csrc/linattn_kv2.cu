// LinearAttention, key side for 64-channel inputs, second generation (round 2): the values are never computed.
// (module_util.py:89-97 PreNorm + :77-86 channel LayerNorm + :170-177 k-softmax / context of the reference.)
//
//   context[h][d][e] = sum_px softmax_px(k)[d] v[e],  v = W_v xn   =>   context = (P^T Xn) W_v^T / S
//
// so the kernel accumulates G[(h,d)][c] = sum_px P[px][(h,d)] xn[px][c] and S[(h,d)] = sum_px P[px][(h,d)] per image, and
// the fold kernel (dac_linattn_fold_g, linattn.cu) applies the constant per-head matrix M_h = W_out,h W_v,h afterwards.
// Against round 1's linattn_kv_kernel<64> per 128-pixel tile: GEMM 1 is N = 128 instead of 256 (k only), the epilogue
// handles 128 instead of 256 accumulator columns and stages P only (no V tile: half the shared-memory stores), GEMM 2 is
// ONE N = 80 MMA per K step (64 channels of the normalised tile itself as the MN-major B operand + 16 columns of ones
// for S) instead of an N = 128 and an N = 16 one.  Tensor pipe ~670 instead of ~1390 cycles per tile, shared-memory
// traffic ~164 instead of ~355 KB per tile - the round-1 kernel kept the shared-memory pipe ~63 % busy with every other
// unit below 45 % (ncu, profiles/).
//
// PreNorm runs INSIDE the kernel: four warps normalise each 128 x 64 tile in place in its pipeline stage (thread = pixel
// row: 8 x LDS.128 in a conflict-free piece order, fp32 mean / centred variance, 8 x STS.128 of the bf16-rounded row -
// exactly what the separate LayerNorm pass wrote to memory), so the 56 us `prenorm` pass of round 1 (134 MB read + 134 MB
// written at 256^2, batch 16) and the normalised tensor are gone.  The same stage is then the K-major A operand of GEMM 1
// and the MN-major B operand of GEMM 2.
//
//   xn = LayerNorm_c(x) (no gain: folded into W_k)     normaliser warps, in place in the pipeline stage
//   k = W_k xn                                          GEMM 1 per 128-pixel tile: M128 x N128 x K64 (two TMEM stages)
//   P = exp(k - c_h)                                    epilogue group h = head h: 32 columns per row, bf16 [pixel][channel]
//   G | S += P^T [xn | 1]                               GEMM 2: M128 x N80 x K128(pixels), both operands MN-major
// TMEM: [0,128) / [128,256) GEMM-1 accumulators, [256,336) / [384,464) {G, S} of alternating images.
// Roles (22 warps): warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 normaliser, warps 6-21 epilogue.
#include <cuda.h>
#include <cuda_runtime.h>
#include <new>

#include "../../include/dac_b200.h"
#include "common.h"
#include "linattn_kv_common.h"
#include "tensormap.h"
#include "tile_common.cuh"

namespace dac {

constexpr uint32_t kKv2Slab = kTileM * 128;     // 128 rows x 64 bf16 (16 KB)
constexpr uint32_t kKv2ColG = 256, kKv2GN = 80, kKv2GStride = 128;   // {G[64], S[16]} per image buffer
constexpr int kKv2Threads = 704;
constexpr int kKv2Stages = 8;

__global__ void __launch_bounds__(kKv2Threads, 1)
linattn_kv2_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW,
                   const __grid_constant__ Kv2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wres = smem;                                         // W_k: 128 rows x 128 B
  uint8_t* pbuf = wres + kKv2Slab;                              // [2 buffers][2 slabs]: P tile, channels 0-63 | 64-127
  uint8_t* ring = pbuf + 4 * kKv2Slab;                          // [kKv2Stages] activation tiles (normalised in place)
  uint8_t* ones = ring + kKv2Stages * kKv2Slab;                 // 128 x 64 bf16 of 1.0 - BEHIND the ring (second B atom)
  uint64_t* bars = reinterpret_cast<uint64_t*>(ones + kKv2Slab);
  uint64_t* full = bars;                 // [8]  TMA landed
  uint64_t* normed = bars + 8;           // [8]  tile normalised (count 128)
  uint64_t* empty = bars + 16;           // [8]  GEMM 2 has read the tile
  uint64_t* acc_full = bars + 24;        // [2]
  uint64_t* acc_empty = bars + 26;       // [2]  count 512 (all four heads)
  uint64_t* p_full = bars + 28;          // [2]  count 512
  uint64_t* p_free = bars + 30;          // [2]
  uint64_t* g_done = bars + 32;          // [2]  the image's {G, S} is complete
  uint64_t* g_flushed = bars + 34;       // [2]  count 128
  uint64_t* w_full = bars + 36;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 37);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int begin, end;
  tile_range(p.tiles, begin, end);
  const int n = end - begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapX);
    tma_prefetch_desc(&mapW);
    for (int s = 0; s < kKv2Stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&normed[s], 128);
      mbar_init(&empty[s], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&acc_full[g], 1);
      mbar_init(&acc_empty[g], 512);
      mbar_init(&p_full[g], 512);
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
  // the tile of ones (second atom of the GEMM-2 B operand: S += P^T 1): every element equal, so the swizzle does not matter
  for (uint32_t i = threadIdx.x; i < kKv2Slab / 16; i += blockDim.x)
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
      mbar_arrive_expect_tx(w_full, kKv2Slab);
      tma_load_2d(wres, &mapW, w_full, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < n; ++i) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], kKv2Slab);
        tma_load_2d(ring + static_cast<size_t>(stage) * kKv2Slab, &mapX, &full[stage], 0, (begin + i) * kTileM);
        if (++stage == kKv2Stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: k(0), k(1); then per tile i: g(i), k(i+2) =====================
    const uint32_t idesc_k = make_idesc_bf16(kTileM, 128);
    const uint32_t idesc_g = make_idesc_bf16(kTileM, kKv2GN) | (1u << 15) | (1u << 16);   // A and B MN-major
    const uint64_t desc_k = make_sw128_desc(0);
    const uint64_t desc_nolbo = desc_k & ~(static_cast<uint64_t>(0x3FFF) << 16);
    // MN-major P operand spanning two 64-channel swizzle atoms: leading byte offset = distance between the slabs
    const uint64_t desc_p = desc_nolbo | (static_cast<uint64_t>(kKv2Slab >> 4) << 16);
    const uint32_t ring_lo = (smem_u32(ring) & 0x3FFFF) >> 4, w_lo = (smem_u32(wres) & 0x3FFFF) >> 4,
                   p_lo0 = (smem_u32(pbuf) & 0x3FFFF) >> 4, ones_lo = (smem_u32(ones) & 0x3FFFF) >> 4,
                   slab_lo = kKv2Slab >> 4;
    int kstage = 0, gstage = 0;
    uint32_t kphase = 0;
    int cur_img = -1, gb = 1;
    uint32_t gacc = 0;
    uint32_t nimg = 0;                       // images started by this CTA
    auto kgemm = [&](int i) {
      mbar_wait(&normed[kstage], kphase);
      mbar_wait(&acc_empty[i & 1], ((i >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint64_t adesc = desc_k | (ring_lo + kstage * slab_lo);
      const uint64_t bdesc = desc_k | w_lo;
      const uint32_t d = tmem_base + (i & 1) * 128;
      if (elect_one()) {
        umma_bf16(d, adesc, bdesc, idesc_k, 0u);
        umma_bf16(d, adesc + 2, bdesc + 2, idesc_k, 1u);
        umma_bf16(d, adesc + 4, bdesc + 4, idesc_k, 1u);
        umma_bf16(d, adesc + 6, bdesc + 6, idesc_k, 1u);
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
        if (cur_img >= 0) {                  // the finished image's {G, S} goes to the flushing group
          if (elect_one()) umma_commit(&g_done[gb]);
          __syncwarp();
        }
        cur_img = img;
        gb ^= 1;
        if (nimg >= 2) {                     // the accumulator's previous image must have been read out
          mbar_wait(&g_flushed[gb], ((nimg >> 1) - 1) & 1);
          tc_fence_after();
        }
        ++nimg;
        gacc = 0;
      }
      const int b = i & 1;
      mbar_wait(&p_full[b], (i >> 1) & 1);
      tc_fence_after();
      const uint32_t a_lo = p_lo0 + b * 2 * slab_lo;
      const uint32_t x_lo = ring_lo + gstage * slab_lo;
      // B = [normalised tile | ones]: the second 64-column atom (16 columns used) sits (ones - stage) bytes further on
      const uint64_t desc_b = desc_nolbo | (static_cast<uint64_t>(ones_lo - x_lo) << 16);
      const uint32_t d = tmem_base + kKv2ColG + gb * kKv2GStride;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)       // 16 pixels (rows of the [pixel][channel] tiles) per K step: +2048 B
          umma_bf16(d, desc_p | (a_lo + ks * 128), desc_b | (x_lo + ks * 128), idesc_g, gacc | (ks ? 1u : 0u));
        umma_commit(&p_free[b]);
        umma_commit(&empty[gstage]);
      }
      __syncwarp();
      gacc = 1;
      if (++gstage == kKv2Stages) gstage = 0;
    };
    mbar_wait(w_full, 0);
    if (n > 0) kgemm(0);
    if (n > 1) kgemm(1);
    for (int i = 0; i < n; ++i) {
      ggemm(i);
      if (i + 2 < n) kgemm(i + 2);
    }
    if (n > 0) {
      if (elect_one()) umma_commit(&g_done[gb]);
      __syncwarp();
    }
  } else if (warp < 6) {
    // ===================== normaliser: thread = pixel row of the tile, in place =====================
    // Row r of a 128B-swizzled tile is the 128 bytes at r * 128 with its 16-byte pieces permuted (piece j at (j ^ (r & 7))):
    // the statistics do not care about the order and the result goes back where it came from.  Step j touches piece
    // j ^ (r & 7), so the eight rows of a quarter-warp hit eight different bank groups (no conflicts).
    const int row = (warp - 2) * 32 + lane;
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < n; ++i) {
      mbar_wait(&full[stage], phase);
      const uint32_t base = smem_u32(ring) + stage * kKv2Slab + row * 128;
      uint64_t x[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t a, b, c, d;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(base + 16 * (j ^ (row & 7))));
        float2 t;
        t = unpack_bf16(a); x[4 * j] = pack_f32x2(t.x, t.y);
        t = unpack_bf16(b); x[4 * j + 1] = pack_f32x2(t.x, t.y);
        t = unpack_bf16(c); x[4 * j + 2] = pack_f32x2(t.x, t.y);
        t = unpack_bf16(d); x[4 * j + 3] = pack_f32x2(t.x, t.y);
      }
      uint64_t s2[4] = {x[0], x[1], x[2], x[3]};
#pragma unroll
      for (int j = 4; j < 32; ++j) s2[j & 3] = add_f32x2(s2[j & 3], x[j]);
      float sa, sb;
      unpack_f32x2(add_f32x2(add_f32x2(s2[0], s2[1]), add_f32x2(s2[2], s2[3])), sa, sb);
      const float mean = (sa + sb) * (1.0f / 64.0f);
      const uint64_t nm = pack_f32x2(-mean, -mean);
      uint64_t q2[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        x[j] = add_f32x2(x[j], nm);
        q2[j & 3] = fma_f32x2(x[j], x[j], q2[j & 3]);
      }
      unpack_f32x2(add_f32x2(add_f32x2(q2[0], q2[1]), add_f32x2(q2[2], q2[3])), sa, sb);
      const float rstd = rsqrtf((sa + sb) * (1.0f / 64.0f) + p.ln_eps);
      const uint64_t r2 = pack_f32x2(rstd, rstd);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float lo, hi;
          unpack_f32x2(mul_f32x2(x[4 * j + q], r2), lo, hi);
          o[q] = pack_bf16(lo, hi);
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + 16 * (j ^ (row & 7))), "r"(o[0]), "r"(o[1]),
                     "r"(o[2]), "r"(o[3])
                     : "memory");
      }
      fence_proxy_async();                   // generic-proxy writes -> visible to the tensor core
      mbar_arrive(&normed[stage]);
      if (++stage == kKv2Stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue: group h = head h, thread = pixel row =====================
    const int quad = warp & 3;
    const int h = (warp - 6) >> 2;           // 0..3
    const int row = quad * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const float shj = p.shift_max[h];
    // The shift only has to keep exp() in range: softmax over the pixels of channel d is invariant to ANY per-channel
    // constant (G and S carry the same factor), so ONE scalar - the largest bound of the head - serves all 32 channels.
    const uint64_t l2e = pack_f32x2(1.4426950408889634f, 1.4426950408889634f), msh = pack_f32x2(-shj, -shj);
    uint32_t nflush = 0;
    // group 0 writes the finished {G, S} of image `img`: TMEM lane = (head, d) -> this warp (quad) holds head `quad`
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
    int prev_img = -1;
    for (int i = 0; i < n; ++i) {
      const int img = (begin + i) / p.tiles_per_image;
      const int b = i & 1;
      const uint32_t pt = smem_u32(pbuf) + static_cast<uint32_t>(b) * 2 * kKv2Slab;
      mbar_wait(&acc_full[b], (i >> 1) & 1);
      tc_fence_after();
      uint32_t rk[32];
      tmem_ld32(lane_base + b * 128 + 32 * h, rk);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&acc_empty[b]);                                            // the accumulator is in registers
      float v[32];
#pragma unroll
      for (int q = 0; q < 32; q += 2) {
        float a0, a1;
        unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(rk[q]), __uint_as_float(rk[q + 1])), l2e, msh), a0, a1);
        v[q] = ex2_approx(a0);
        v[q + 1] = ex2_approx(a1);
      }
      mbar_wait(&p_free[b], ((i >> 1) & 1) ^ 1);                             // GEMM 2 of the buffer's previous tile is done
      chunk_stage_bf16_s(pt, row, 32 * h, v);
      fence_proxy_async();
      mbar_arrive(&p_full[b]);
      if (h == 0 && prev_img >= 0 && img != prev_img) flush(prev_img);
      prev_img = img;
    }
    if (h == 0 && n > 0) flush(prev_img);
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

int dac_kv2_smem_bytes() { return (1 + 4 + kKv2Stages + 1) * (int)kKv2Slab + 1024 + 512; }

int dac_kv2_launch(const CUtensorMap& mapX, const CUtensorMap& mapW, const Kv2Params& kp, int grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(linattn_kv2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dac_kv2_smem_bytes());
    if (e != cudaSuccess) return set_error(-12, "dac_linattn_kv: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  linattn_kv2_kernel<<<grid, kKv2Threads, dac_kv2_smem_bytes(), st>>>(mapX, mapW, kp);
  return check_launch("linattn_kv2_kernel");
}
