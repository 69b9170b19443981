// Host side of the tcgen05 implicit-GEMM convolution: plan creation (validation, TMA descriptor encoding, tile /
// pipeline sizing, epilogue-kernel selection) and launch.  The kernel itself is in conv_kernel.cuh.
// Replaces nn.Conv2d / nn.Linear + the pointwise ops around them in module_util.py:100-153,157-185 and
// attention.py:37-64,152-261 of the reference.
#include <cuda.h>
#include <cuda_runtime.h>
#include <mutex>
#include <new>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/dac_b200.h"
#include "common.h"
#include "conv_kernel.cuh"
#include "tensormap.h"

namespace dac {

// ------------------------------------------------------------------------------------------- host side
}  // namespace dac

struct dac_conv_plan {
  CUtensorMap mapA0, mapA1, mapW, mapOut, mapOut2, mapR0, mapR1, mapWR, mapRes;
  dac::ConvKParams kp;
  dac::ConvKernelFn kernel;
  int grid;
  int smem;
  int tiles;
};

using namespace dac;

static int encode_act_map(CUtensorMap* m, const void* ptr, int c, int ld, int W, int H, int B, int tile_w,
                          int box_rows, int stride) {   // tile_w = box width in pixels
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return set_error(-10, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {(cuuint32_t)kChunkK, (cuuint32_t)(tile_w * stride), (cuuint32_t)(box_rows * stride), 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(-11, "cuTensorMapEncodeTiled(activation) failed: CUresult %d", (int)r);
  return 0;
}

// CTA-pair mode on by default?  (DAC_CTA2 overrides either way.)
static const bool kCta2Default = true;
// ... and for layers with resident non-pair weights (DAC_CTA2_RES=0 / 1 overrides)
static const bool kCta2ResDefault = true;
static bool cta2_res_on() {
  const char* e = getenv("DAC_CTA2_RES");
  return e ? atoi(e) != 0 : kCta2ResDefault;
}

extern "C" int dac_conv_create(const dac_conv_desc* d, dac_conv_t* out) {
  if (!d || !out) return set_error(-1, "dac_conv_create: null argument");
  *out = nullptr;
  if (d->tile_h * d->tile_w != kTileM || d->tile_w < 8 || (d->tile_w & (d->tile_w - 1)))
    return set_error(-2, "dac_conv_create: tile %dx%d must have 128 pixels, width multiple of 8", d->tile_h, d->tile_w);
  if (d->c0 <= 0 || d->c0 % kChunkK || d->c1 % kChunkK || d->ld0 % 8 || (d->c1 && d->ld1 % 8))
    return set_error(-2, "dac_conv_create: channels (%d,%d) must be multiples of 64, pitches of 8", d->c0, d->c1);
  if (d->block_n < 16 || d->block_n > 256 || d->block_n % 16 || d->cout_pad % d->block_n || d->cout > d->cout_pad)
    return set_error(-2, "dac_conv_create: bad block_n %d / cout %d / cout_pad %d", d->block_n, d->cout, d->cout_pad);
  if (d->ntaps < 1 || d->ntaps > 16 || (d->ngroups != 1 && d->ngroups != 4) || (d->stride != 1 && d->stride != 2))
    return set_error(-2, "dac_conv_create: bad ntaps/ngroups/stride");
  if (d->ndy < 1 || d->ncols < 1 || d->ndy * d->ncols != d->ntaps || (d->ndy > 1 && d->stride != 1))
    return set_error(-2, "dac_conv_create: column groups (%d x %d) must cover the %d taps; ndy > 1 needs stride 1",
                     d->ncols, d->ndy, d->ntaps);
  // halo loads: a k x k window of taps (3 x 3, or the 2 x 2 of one parity group of the folded upsample conv) served by
  // ONE (tile_h + k - 1) x (tile_w + k - 1) pixel box per K chunk; tap i = (i / k, i % k) pixels into the box
  const int halo_k = !d->halo ? 0 : (d->ntaps == 9 ? 3 : (d->ntaps == 4 ? 2 : -1));
  if (d->halo && (d->stride != 1 || halo_k < 0 || d->ndy != d->ntaps || d->ncols != 1 || d->tile_w != 8 ||
                  (d->ngroups != 1 && halo_k != 2) || d->halo > 2))
    return set_error(-2, "dac_conv_create: halo loads need a 3x3 (or per-group 2x2) stride-1 conv, tile_w 8, ncols 1, ndy = ntaps");
  const int a_rows = d->halo ? d->tile_h + halo_k - 1 : (d->stride == 1 ? d->tile_h + d->ndy - 1 : d->tile_h);
  const int box_w = d->halo ? d->tile_w + halo_k - 1 : d->tile_w;           // rows x box_w pixels land in smem
  if (d->tile_w * d->stride > 256 || a_rows * d->stride > 256)
    return set_error(-2, "dac_conv_create: TMA box exceeds 256");
  if ((reinterpret_cast<uintptr_t>(d->src0) | reinterpret_cast<uintptr_t>(d->src1) |
       reinterpret_cast<uintptr_t>(d->weight) | reinterpret_cast<uintptr_t>(d->out) |
       reinterpret_cast<uintptr_t>(d->res) | reinterpret_cast<uintptr_t>(d->res2)) & 15)
    return set_error(-2, "dac_conv_create: pointers must be 16-byte aligned");
  if (!d->out && !d->out_nchw && !d->out_f32 && d->epi != DAC_EPI_KVCTX)
    return set_error(-2, "dac_conv_create: no output");
  if ((d->out_f32 && (d->out_f32_ld & 3)) || (d->res_f32 && (d->res_f32_ld & 3)) ||
      ((reinterpret_cast<uintptr_t>(d->out_f32) | reinterpret_cast<uintptr_t>(d->res_f32)) & 15))
    return set_error(-2, "dac_conv_create: fp32 stream pointers/pitches must be 16-byte aligned");
  if ((d->out_f32 || d->res_f32) && d->epi != DAC_EPI_PLAIN)
    return set_error(-2, "dac_conv_create: fp32 stream only with the PLAIN epilogue");
  const bool nchw = d->out_nchw != nullptr;
  if (nchw) {
    if (d->block_n != (d->pair ? 32 : 16) || d->out_nchw_c > 16 || d->epi != DAC_EPI_PLAIN || d->act != DAC_ACT_NONE || d->film ||
        d->out || d->out_f32)
      return set_error(-2, "dac_conv_create: fp32 NCHW output needs block_n 16, <= 16 channels, plain epilogue");
  } else if (d->epi == DAC_EPI_LN) {
    if (d->cout != d->block_n || d->cout_pad != d->cout || !d->ln_g || (d->cout & 63) || !d->out)
      return set_error(-2, "dac_conv_create: LN epilogue needs a single N tile with cout %% 64 == 0");
  } else if (d->epi == DAC_EPI_QKV) {
    const bool q_only = d->cout == 128;   // k | v go through a KVCTX plan instead
    if (d->block_n != 128 || (d->cout != 384 && !q_only) || !d->out || (!q_only && !d->out_planar) ||
        d->out_scale > 1 || (reinterpret_cast<uintptr_t>(d->out_planar) & 127))
      return set_error(-2, "dac_conv_create: QKV epilogue needs block_n 128, cout 384 (128: q only), out (q) and "
                           "out_planar (k|v)");
  } else if (d->epi == DAC_EPI_KVCTX) {
    if (d->block_n != 256 || d->cout != 256 || d->cout_pad != 256 || !d->kv_shift || !d->ctx_acc || d->out ||
        d->out_f32 || d->ngroups != 1 || d->per_image_w || d->out_scale > 1 ||
        ((reinterpret_cast<uintptr_t>(d->kv_shift) | reinterpret_cast<uintptr_t>(d->ctx_acc)) & 15))
      return set_error(-2, "dac_conv_create: KVCTX epilogue needs block_n = cout = 256, kv_shift, ctx_acc, no out");
  } else if (d->epi == DAC_EPI_GEGLU) {
    if (!d->bias || (d->block_n & 127) || d->cout != d->cout_pad || !d->out)
      return set_error(-2, "dac_conv_create: GEGLU epilogue needs bias and block_n %% 128 == 0");
  } else if (d->epi == DAC_EPI_PLAIN) {
    if ((d->block_n & 63) || (d->cout & 31))
      return set_error(-2, "dac_conv_create: plain epilogue needs block_n %% 64 == 0 and cout %% 32 == 0 (got %d, %d)",
                       d->block_n, d->cout);
  } else {
    return set_error(-2, "dac_conv_create: unknown epilogue %d", d->epi);
  }
  if (d->film && ((d->film_ld | d->film_off) & 3)) return set_error(-2, "dac_conv_create: film_ld/film_off %% 4");
  if ((reinterpret_cast<uintptr_t>(d->bias) | reinterpret_cast<uintptr_t>(d->bias_img) |
       reinterpret_cast<uintptr_t>(d->film) | reinterpret_cast<uintptr_t>(d->ln_g)) & 15)
    return set_error(-2, "dac_conv_create: parameter vectors must be 16-byte aligned");
  if (d->stats_out && (d->epi != DAC_EPI_PLAIN || nchw || d->cout_pad != d->block_n || d->out_scale > 1 ||
                       (reinterpret_cast<uintptr_t>(d->stats_out) & 7)))
    return set_error(-2, "dac_conv_create: stats_out needs the PLAIN epilogue and a single N tile");
  if ((d->ln_stats != nullptr) != (d->ln_colsum != nullptr) || (d->ln_stats && d->epi != DAC_EPI_QKV && d->epi != DAC_EPI_KVCTX) ||
      ((reinterpret_cast<uintptr_t>(d->ln_stats) & 7) | (reinterpret_cast<uintptr_t>(d->ln_colsum) & 15)))
    return set_error(-2, "dac_conv_create: ln_stats / ln_colsum go together, QKV and KVCTX epilogues only");
  const bool fused_res = d->rsrc0 != nullptr;
  if (fused_res) {
    if (d->epi != DAC_EPI_PLAIN || nchw || d->cout_pad != d->block_n || 2 * d->block_n > 256 || !d->rweight ||
        d->rc0 <= 0 || d->rc0 % kChunkK || d->rc1 % kChunkK || (d->rld0 & 7) || (d->rsrc1 && (d->rld1 & 7)) ||
        d->stride != 1 || d->ngroups != 1 || d->per_image_w ||
        ((reinterpret_cast<uintptr_t>(d->rsrc0) | reinterpret_cast<uintptr_t>(d->rsrc1) |
          reinterpret_cast<uintptr_t>(d->rweight)) & 15))
      return set_error(-2, "dac_conv_create: fused skip conv needs PLAIN epilogue, one N tile <= 128, stride 1");
  }
  if (d->pair) {
    const bool pair_ok = nchw ? (d->block_n == 32 && d->cout_pad == 32 && !d->rsrc0)
                              : (d->block_n == 128 && d->cout == 128 && d->cout_pad == 128);
    if (!d->halo || !pair_ok || d->epi != DAC_EPI_PLAIN ||
        d->stats_out || d->per_image_w || d->rc0 % 128 || d->rc1 % 128 || d->c0 % 128 || d->c1 % 128 || d->out_scale > 1 || d->out_f32 ||
        d->res_f32 || d->bias_img)
      return set_error(-2, "dac_conv_create: pixel-pair mode needs a haloed 3x3 conv, PLAIN epilogue, wide views with "
                           "block_n = cout = 128 and channel counts that are multiples of 128");
  }
  // the fp32 residual stream of the ViT blocks has its own flavour (residual chunks prefetched two ahead)
  const bool f32_stream = d->res_f32 && !d->res && !d->res2 && !d->bias_img && !d->stats_out && !d->rsrc0;
  ConvKernelFn kernel = pick_conv_kernel(d->epi, d->act, d->film != nullptr, nchw, f32_stream);
  if (!kernel) return set_error(-2, "dac_conv_create: unsupported activation / FiLM combination (%d, %d)", d->act,
                                d->film != nullptr);
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
  k.tile_w_shift = 0;
  while ((1 << k.tile_w_shift) < d->tile_w) ++k.tile_w_shift;
  k.n_tiles = d->cout_pad / d->block_n;
  k.ngroups = d->ngroups; k.ntaps = d->ntaps;
  k.fd_ntiles = make_fast_div(k.n_tiles); k.fd_mtiles = make_fast_div(k.m_tiles);
  k.fd_mpairs = make_fast_div(k.m_tiles > 1 ? k.m_tiles / 2 : 1);
  k.fd_tx = make_fast_div(k.tiles_x); k.fd_ty = make_fast_div(k.tiles_y);
  k.chunks0 = d->c0 / kChunkK; k.chunks1 = d->c1 / kChunkK; k.c0 = d->c0;
  k.per_image_w = d->per_image_w;
  k.b_bytes = (uint32_t)d->block_n * kChunkK * 2;
  k.ndy = d->ndy; k.ncols = d->ncols;
  memcpy(k.col_dx, d->col_dx, sizeof(k.col_dx));
  memcpy(k.col_dy0, d->col_dy0, sizeof(k.col_dy0));
  memcpy(k.col_tap, d->col_tap, sizeof(k.col_tap));
  k.r_chunks0 = fused_res ? d->rc0 / kChunkK : 0;
  k.r_chunks1 = fused_res ? d->rc1 / kChunkK : 0;
  k.r_a_bytes = (uint32_t)d->tile_h * d->tile_w * kChunkK * 2;
  k.r_b_bytes = d->pair ? 64u * kChunkK * 2 : k.b_bytes;
  k.a_bytes = (uint32_t)a_rows * box_w * kChunkK * 2;
  k.a_slot = (k.a_bytes + 1023u) & ~1023u;
  k.a_sbo = d->halo ? (uint32_t)box_w * 128u : 1024u;          // halo, tile_w 8: one (padded) pixel row per 8-row group
  k.halo = d->halo;
  for (int i = 0; i < d->ndy; ++i) {
    // tap i of a load: halo = (ky, kx) = (i / k, i % k) pixels into the haloed box; else i tile rows down
    const uint32_t off = d->halo ? (uint32_t)((i / halo_k) * box_w + (i % halo_k)) * 128u : (uint32_t)i * d->tile_w * 128u;
    k.tap_off[i] = (uint16_t)(off >> 4);
  }
  k.cout = d->cout;
  pl->kernel = kernel;
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
  k.out_planar = static_cast<__nv_bfloat16*>(d->out_planar);
  k.stats_out = d->stats_out; k.stats_eps = d->stats_eps;
  k.ln_stats = d->ln_stats; k.ln_colsum = d->ln_colsum;
  k.out_nchw = d->out_nchw; k.nchw_c = d->out_nchw_c; k.nchw_h = d->out_nchw_h; k.nchw_w = d->out_nchw_w;

  // Weights stay resident in shared memory when the whole tensor fits beside >= 3 activation stages: the
  // mainloop then streams activations only (L2 -> SM ingest is what bounds the 64/128-channel layers).
  const int smem_budget = 227 * 1024 - 1024 /*align*/ - 256 /*barriers*/ - 4096 /*FiLM stage, 2 groups*/;
  const int chunks = k.chunks0 + k.chunks1;
  const long long res_bytes = (long long)k.n_tiles * d->ntaps * chunks * k.b_bytes;
  const long long kv_extra = d->epi == DAC_EPI_KVCTX ? 2ll * kKvStageBytes : 0;   // P / V head tiles, both groups
  // (per parity group: tiles are numbered group-major and the kernel swaps the resident set when the group changes)
  bool resident = !d->per_image_w && res_bytes <= 160 * 1024 &&
                  (smem_budget - res_bytes - kv_extra) / (long long)k.a_slot >= 3;
  k.b_res_bytes = resident ? (uint32_t)res_bytes : 0u;
  k.pair = d->pair ? 1 : 0;
  k.bias_sh = (d->bias && !d->film && !nchw && !d->pair && d->block_n <= 512 && !getenv("DAC_NO_BIAS_SH") &&
               (d->epi == DAC_EPI_PLAIN || d->epi == DAC_EPI_GEGLU)) ? 1 : 0;
  // CTA-pair mode (conv_kernel.cuh): streamed-weight layers on 2-CTA clusters, each CTA loading half of every weight tile.
  // DAC_CTA2=0 switches it off (A/B runs); layers whose weights are resident, the pixel-pair / fused-skip / per-image-weight
  // / KVCTX / QKV / NCHW flavours and odd M-tile counts keep the 1-CTA kernel.
  const char* cta2_env = getenv("DAC_CTA2");
  const bool cta2_on = cta2_env ? atoi(cta2_env) != 0 : kCta2Default;
  // (pixel-pair layers: only when the weight tensor carries the per-rank layouts, dac_conv_desc.pair == 2)
  // (resident non-pair weights - haloed 3x3, parity-group upsample, stride-2 and 7-tap stem layers: each CTA keeps half of the
  // rows of every weight tile; DAC_CTA2_RES=0 / 1 selects it)
  const bool cta2 = cta2_on && !(fused_res && (getenv("DAC_NO_CTA2_SKIP") || resident)) && !d->per_image_w && !nchw &&
                    d->epi != DAC_EPI_KVCTX && d->epi != DAC_EPI_QKV && (k.m_tiles % 2) == 0 && (d->block_n % 32) == 0 &&
                    d->block_n >= 64 && !d->stats_out &&
                    (d->pair ? (d->pair == 2 && !getenv("DAC_NO_CTA2_PAIR"))
                             : (d->ngroups == 1 && !resident && !d->halo) || ((resident || d->halo) && cta2_res_on()));
  if (cta2) {
    ConvKernelFn kernel2 = pick_conv_kernel(d->epi, d->act, d->film != nullptr, nchw, f32_stream, true);
    if (kernel2) {
      kernel = kernel2;
      pl->kernel = kernel2;
      k.cta2 = 1;
      if (!d->pair) {
        k.b_bytes >>= 1;               // per CTA: half of the block_n weight rows of a K step
        k.b_res_bytes >>= 1;           // ... and of the resident weight tensor
      }
      k.r_b_bytes >>= 1;               // ... and half of the fused skip conv's weight rows
    }
  }
  if (d->pair) {   // one 192-row block per (64-channel source slice, ky), always resident
    k.b_res_bytes = (uint32_t)(chunks / 2) * 3u * (3u * (uint32_t)(d->block_n / 2) * 128u);
    k.film_cols = d->block_n / 2;
    resident = true;
    if ((smem_budget - (long long)k.b_res_bytes) / (long long)k.a_slot < 2) {
      delete pl;
      return set_error(-2, "dac_conv_create: pixel-pair weights (%u B) leave no room for two pipeline stages", k.b_res_bytes);
    }
  }
  uint32_t stage_bytes = k.a_slot + (resident ? 0u : (uint32_t)d->ndy * k.b_bytes);
  if (fused_res && stage_bytes < k.r_a_bytes + k.r_b_bytes) stage_bytes = k.r_a_bytes + k.r_b_bytes;
  // bf16 NHWC output through a swizzled staging tile + TMA store (coalesced, clipped by the tensor map) whenever the
  // staging tile leaves room for >= 3 pipeline stages; otherwise each thread stores its own row directly.
  const int out_cols = d->epi == DAC_EPI_GEGLU ? d->block_n / 2 : d->block_n;
  uint32_t stg_bytes = 0;
  // (the planar k|v map of the QKV epilogue needs 16-byte row pitches: OW % 8 == 0, else direct stores)
  if (d->out && !nchw && k.out_scale == 1 && out_cols % 64 == 0 && d->cout % 64 == 0 && (d->out_coff % 64) == 0 &&
      !(d->epi == DAC_EPI_QKV && (d->OW & 7)) && !getenv("DAC_NO_TMA_STORE")) {
    const uint32_t want = (uint32_t)(out_cols / 64) * kTileM * 128;     // per epilogue group; two groups
    int without = (smem_budget - (int)k.b_res_bytes) / (int)stage_bytes;
    if (without > kMaxStages) without = kMaxStages;
    const int with = (smem_budget - (int)k.b_res_bytes - 2 * (int)want) / (int)stage_bytes;
    // (pixel-pair mode with a residual: one K chunk keeps the tensor pipe busy for 1344 cycles, three stages are enough,
    // and direct residual loads - one 256-byte row per thread - cost far more than the pipeline depth buys: 124 -> 95 us;
    // without a residual the deeper pipeline wins: FiLM 91 vs 94 us, fused skip 107 vs 130 us)
    // (CTA-pair mode: the leader waits for two TMA streams per stage, a deeper pipeline is worth more than coalesced stores
    // below five stages - ViT c_fc 64.6 us with four stages + staging, 60.5 us with seven stages and direct stores)
    if ((k.cta2 && !d->pair) ? with >= 5 : (with >= 4 || (with >= 3 && (with >= without || (d->pair && d->res))))) stg_bytes = want;
  }
  if (d->epi == DAC_EPI_KVCTX) stg_bytes = kKvStageBytes;   // P / V head tiles of each epilogue group
  k.kv_shift = d->kv_shift; k.ctx_acc = d->ctx_acc;
  if (d->epi == DAC_EPI_KVCTX) {
    // softmax over the pixels of a channel is invariant to any per-channel constant: one scalar shift per head (the
    // largest bound) keeps exp() in range just as well and costs the epilogue no loads
    float sh[128];
    cudaError_t ce = cudaMemcpy(sh, d->kv_shift, sizeof(sh), cudaMemcpyDeviceToHost);
    if (ce != cudaSuccess) {
      delete pl;
      return set_error(-20, "dac_conv_create: reading kv_shift: %s", cudaGetErrorString(ce));
    }
    for (int h = 0; h < 4; ++h) {
      float m = sh[h * 32];
      for (int j = 1; j < 32; ++j) m = sh[h * 32 + j] > m ? sh[h * 32 + j] : m;
      k.kv_shift_max[h] = m;
    }
  }
  k.ctx_slots = d->ctx_slots; k.ctx_tpi = k.tiles_x * k.tiles_y;
  // FiLM parameters in TMEM when three block_n-wide regions fit one accumulator stage (the 64-channel layers, which
  // are the shared-memory-bound ones); alignment of the float4 parameter loads needs cout % 4 == 0 (validated above)
  k.film_tmem = (d->film && !fused_res && d->block_n % 32 == 0 &&
                 d->block_n + 2 * (k.film_cols ? k.film_cols : d->block_n) <= (int)kAccStride &&
                 !getenv("DAC_NO_FILM_TMEM")) ? 1 : 0;
  k.stg_bytes = stg_bytes;
  k.dbg = getenv("DAC_EPI_DEBUG") ? atoi(getenv("DAC_EPI_DEBUG")) : 0;   // profiling only: wrong results by design
  k.stg_count = stg_bytes ? 2 : 1;   // one staging tile per epilogue group
  int stages = (smem_budget - (int)k.b_res_bytes - (int)stg_bytes * k.stg_count) / (int)stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) { delete pl; return set_error(-2, "dac_conv_create: tile does not fit shared memory"); }
  k.stages = stages;
  pl->smem = (int)k.b_res_bytes + stages * (int)stage_bytes + (int)stg_bytes * k.stg_count + 1024 + 256 + 4096;
  pl->tiles = k.ngroups * k.m_tiles * k.n_tiles;

  int rc = encode_act_map(&pl->mapA0, d->src0, d->c0, d->ld0, d->W, d->H, d->B, box_w, a_rows, d->stride);
  if (rc == 0) {
    if (d->c1 > 0)
      rc = encode_act_map(&pl->mapA1, d->src1, d->c1, d->ld1, d->W, d->H, d->B, box_w, a_rows, d->stride);
    else
      pl->mapA1 = pl->mapA0;
  }
  if (rc == 0) {
    PFN_encodeTiled enc = get_encode_fn();
    const int ctot = d->pair ? (d->c0 + d->c1) / 2 : d->c0 + d->c1;
    const long long Z = d->pair ? (d->pair == 2 ? 9 : 3) : (long long)d->ngroups * d->ntaps * (d->per_image_w ? d->B : 1);
    const int wrows = d->pair ? 3 * (d->block_n / 2) : d->cout_pad;
    cuuint64_t dims[3] = {(cuuint64_t)ctot, (cuuint64_t)wrows, (cuuint64_t)Z};
    cuuint64_t strides[2] = {(cuuint64_t)ctot * 2, (cuuint64_t)wrows * ctot * 2};
    cuuint32_t box[3] = {(cuuint32_t)kChunkK, (cuuint32_t)(d->pair ? wrows : (k.cta2 ? d->block_n / 2 : d->block_n)), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&pl->mapW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(d->weight), dims, strides,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) rc = set_error(-11, "cuTensorMapEncodeTiled(weight) failed: CUresult %d", (int)r);
  }
  if (rc == 0 && stg_bytes && d->epi != DAC_EPI_KVCTX) {
    PFN_encodeTiled enc = get_encode_fn();
    const int valid_c = d->epi == DAC_EPI_GEGLU ? d->cout / 2 : (d->epi == DAC_EPI_QKV ? 128 : d->cout);
    cuuint64_t dims[4] = {(cuuint64_t)(d->out_coff + valid_c), (cuuint64_t)k.OWf, (cuuint64_t)k.OHf, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->out_ld * 2, (cuuint64_t)k.OWf * d->out_ld * 2,
                             (cuuint64_t)k.OHf * k.OWf * d->out_ld * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)d->tile_w, (cuuint32_t)d->tile_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&pl->mapOut, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d->out, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) rc = set_error(-11, "cuTensorMapEncodeTiled(output) failed: CUresult %d", (int)r);
    pl->mapOut2 = pl->mapOut;
    pl->mapRes = pl->mapOut;
    // bf16 residual through TMA: same box as the output tile, landed in the staging tile before the epilogue runs
    if (rc == 0 && d->res && (d->epi == DAC_EPI_PLAIN || d->epi == DAC_EPI_LN) && d->cout == d->cout_pad &&
        !getenv("DAC_NO_TMA_RES")) {
      cuuint64_t rd[4] = {(cuuint64_t)d->cout, (cuuint64_t)k.OWf, (cuuint64_t)k.OHf, (cuuint64_t)d->B};
      cuuint64_t rs[3] = {(cuuint64_t)d->res_ld * 2, (cuuint64_t)k.OWf * d->res_ld * 2,
                          (cuuint64_t)k.OHf * k.OWf * d->res_ld * 2};
      r = enc(&pl->mapRes, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d->res), rd, rs, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) rc = set_error(-11, "cuTensorMapEncodeTiled(residual) failed: CUresult %d", (int)r);
      else pl->kp.res_tma = 1;
    }
    if (rc == 0 && d->epi == DAC_EPI_QKV && d->out_planar) {
      // planar k|v tensor [B][256][OH][OW]: box = tile_w x tile_h pixels x 128 channels, un-swizzled, i.e. the
      // [channel][pixel] staging tile the epilogue wrote
      cuuint64_t pd[4] = {(cuuint64_t)d->OW, (cuuint64_t)d->OH, 256, (cuuint64_t)d->B};
      cuuint64_t ps[3] = {(cuuint64_t)d->OW * 2, (cuuint64_t)d->OH * d->OW * 2, (cuuint64_t)256 * d->OH * d->OW * 2};
      cuuint32_t pb[4] = {(cuuint32_t)d->tile_w, (cuuint32_t)d->tile_h, 128, 1};
      r = enc(&pl->mapOut2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d->out_planar, pd, ps, pb, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) rc = set_error(-11, "cuTensorMapEncodeTiled(planar k|v) failed: CUresult %d", (int)r);
    }
  } else if (rc == 0) {
    pl->mapOut = pl->mapA0;
    pl->mapOut2 = pl->mapA0;
    pl->mapRes = pl->mapA0;
  }
  pl->mapR0 = pl->mapA0; pl->mapR1 = pl->mapA0; pl->mapWR = pl->mapW;
  if (rc == 0 && fused_res) {
    rc = encode_act_map(&pl->mapR0, d->rsrc0, d->rc0, d->rld0, d->W, d->H, d->B, d->tile_w, d->tile_h, 1);
    if (rc == 0 && d->rc1 > 0)
      rc = encode_act_map(&pl->mapR1, d->rsrc1, d->rc1, d->rld1, d->W, d->H, d->B, d->tile_w, d->tile_h, 1);
    if (rc == 0) {
      PFN_encodeTiled enc = get_encode_fn();
      // pair mode: the skip weight is the real [64][rc / 2] matrix; every wide chunk multiplies one 64-row tile of it
      const int rct = d->pair ? (d->rc0 + d->rc1) / 2 : d->rc0 + d->rc1;
      const int rrows = d->pair ? 64 : d->cout_pad;
      cuuint64_t dims[3] = {(cuuint64_t)rct, (cuuint64_t)rrows, 1};
      cuuint64_t strides[2] = {(cuuint64_t)rct * 2, (cuuint64_t)rrows * rct * 2};
      cuuint32_t box[3] = {(cuuint32_t)kChunkK, (cuuint32_t)((d->pair ? 64 : d->block_n) / (k.cta2 ? 2 : 1)), 1};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = enc(&pl->mapWR, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(d->rweight), dims, strides,
                       box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) rc = set_error(-11, "cuTensorMapEncodeTiled(skip weight) failed: CUresult %d", (int)r);
    }
  }
  if (rc != 0) { delete pl; return rc; }

  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  pl->grid = pl->tiles < sms ? pl->tiles : sms;
  if (k.cta2) {
    // one cluster = two CTAs on the two SMs of a TPC; as many clusters as are resident at once (a TPC with one usable SM
    // hosts none), each with a contiguous range of pair tiles
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sms & ~1);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = pl->smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int ncl = 0;
    cudaFuncSetAttribute(reinterpret_cast<const void*>(kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (cudaOccupancyMaxActiveClusters(&ncl, reinterpret_cast<const void*>(kernel), &cfg) != cudaSuccess || ncl <= 0) {
      cudaGetLastError();
      delete pl;
      return set_error(-12, "dac_conv_create: no 2-CTA cluster of this kernel fits the device (DAC_CTA2=0 disables the mode)");
    }
    const int pairs = pl->tiles / 2;
    pl->grid = 2 * (pairs < ncl ? pairs : ncl);
    if (getenv("DAC_CTA2_DEBUG"))
      fprintf(stderr, "[cta2] tiles %d pairs %d max active clusters %d grid %d smem %d stages %d block_n %d\n", pl->tiles, pairs,
              ncl, pl->grid, pl->smem, k.stages, k.block_n);
  }
  if (k.ctx_acc) {
    const int need = 2 * max_image_span(d->B, k.ctx_tpi, pl->grid);
    if (k.n_tiles != 1 || k.ngroups != 1 || d->ctx_slots < need) {
      const int have = d->ctx_slots;
      delete pl;
      return set_error(-2, "dac_conv_create: KVCTX needs ctx_slots >= %d (dac_linattn_ctx_slots), got %d", need, have);
    }
  }
  if (cudaFuncSetAttribute(reinterpret_cast<const void*>(kernel), cudaFuncAttributeMaxDynamicSharedMemorySize,
                           227 * 1024) != cudaSuccess) {
    cudaGetLastError();
    delete pl;
    return set_error(-12, "dac_conv_create: cannot raise the dynamic shared memory limit (not an sm_100 device?)");
  }
  *out = pl;
  return 0;
}

extern "C" int32_t dac_linattn_ctx_slots(int32_t B, int32_t tiles_per_image, int32_t groups) {
  if (B <= 0 || tiles_per_image <= 0 || groups <= 0) return 0;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      sms <= 0) {
    cudaGetLastError();
    sms = 148;
  }
  const long long tiles = static_cast<long long>(B) * tiles_per_image;
  const int grid = tiles < sms ? static_cast<int>(tiles) : sms;
  return groups * max_image_span(B, tiles_per_image, grid);
}

extern "C" int dac_conv_launch(dac_conv_t pl, dac_stream_t stream) {
  if (!pl) return set_error(-1, "dac_conv_launch: null plan");
  if (pl->kp.ctx_acc) {   // KVCTX: slots no CTA writes (an image spanning fewer CTAs than the widest one) must read as zero
    cudaError_t e = cudaMemsetAsync(pl->kp.ctx_acc, 0, sizeof(float) * pl->kp.B * 4 * pl->kp.ctx_slots * kCtxRecord,
                                    static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return set_error(-20, "dac_conv_launch: memset failed: %s", cudaGetErrorString(e));
  }
  if (pl->kp.ctx_acc)   // behind a memset node: a plain launch
    pl->kernel<<<pl->grid, kThreads, pl->smem, static_cast<cudaStream_t>(stream)>>>(pl->mapA0, pl->mapA1, pl->mapW,
                                                                                    pl->mapOut, pl->mapOut2, pl->mapR0,
                                                                                    pl->mapR1, pl->mapWR, pl->mapRes, pl->kp);
  else
    launch_k_cluster(pl->kernel, dim3(pl->grid), dim3(kThreads), pl->smem, static_cast<cudaStream_t>(stream), pl->kp.cta2 ? 2 : 1,
             pl->mapA0, pl->mapA1, pl->mapW, pl->mapOut, pl->mapOut2, pl->mapR0, pl->mapR1, pl->mapWR, pl->mapRes, pl->kp);
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
