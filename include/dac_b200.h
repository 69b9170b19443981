/*
 * dac_b200.h - C ABI of libdac_b200.so: the B200 (sm_100a) kernels behind the DA-CLIP
 * universal-restoration inference path.
 *
 * The reference (yeeecheng/DA-CLIP) is pure Python on PyTorch and has no FFI layer of its own;
 * each entry point below replaces the PyTorch op sequence of the reference site it cites
 * (paths relative to /root/reference/universal-image-restoration unless noted):
 *   SDE  = utils/sde_utils.py
 *   ARCH = config/daclip-sde/models/modules/DenoisingUNet_arch.py
 *   MU   = config/daclip-sde/models/modules/module_util.py
 *   ATT  = config/daclip-sde/models/modules/attention.py
 *   TR   = open_clip/transformer.py
 *
 * Conventions: plain pointers and sizes only (no torch types).  All pointers are DEVICE pointers
 * owned by the caller unless the name says `host_`.  Activations are NHWC bf16, SDE state is
 * NCHW fp32 (the reference's layout).  Every call enqueues work on `stream` and returns 0, or a
 * negative code with a message retrievable through dac_last_error().  Nothing here allocates
 * device memory; plan objects own only host memory (TMA descriptors).
 */
#ifndef DAC_B200_H
#define DAC_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* dac_stream_t; /* cudaStream_t */

int dac_version(void);
/* Programmatic dependent launch for the calling thread's following launches (0 = off, the default): kernels of this
 * library that wait with griddepcontrol before their first dependent access are then launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization, so their set-up overlaps the tail of the previous kernel in the
 * stream.  The caller raises it only when the previous operation in the stream is a kernel of this library (the UNet
 * engine does, step by step, while capturing its CUDA graph). */
void dac_set_pdl(int32_t on);
const char* dac_last_error(void);
/* Number of kernels this library has launched since load (or since the last reset). */
int64_t dac_launch_count(void);
void dac_reset_launch_count(void);
/* sizeof(dac_conv_desc) / sizeof(dac_embed_weights) as compiled: lets a binding detect a stale library. */
int dac_abi_sizes(int32_t* conv_desc_bytes, int32_t* embed_weights_bytes);

/* ------------------------------------------------------------------ SDE updates (fp32, HBM-bound)
 * x, mu, net, eps, out: [n] fp32 (any layout, elementwise).  out may alias x.
 * mode 0: reverse-SDE step     SDE:44-45,177-178,183-187   coef = {theta_t, sigma_t^2, sigma_bar_t, dt, sigma_t, sqrt(dt)}
 * mode 1: posterior step       SDE:205-231,245-247         coef = {term1, term2, std, exp(Theta_t dt), sigma_bar_t}
 * mode 2: probability-flow ODE SDE:47-48,180-181           coef = {theta_t, 0.5*sigma_t^2, sigma_bar_t, dt}
 * host_coef: 8 floats on the HOST (read at call time), computed in fp32 exactly as the reference computes its
 * 0-dim tensors; the kernel applies them in the reference's operation order without FMA contraction, so the
 * update is bit-identical to the PyTorch fp32 expression. */
int dac_sde_step(int mode, const float* x, const float* mu, const float* net, const float* eps, float* out,
                 int64_t n, const float* host_coef, dac_stream_t stream);
/* The same updates inside a device-driven sampling loop (SDE:297-313: the host loop over t): every step of the loop is
 * the same sequence of launches, so it can be one CUDA graph replayed T times with no host work in between.
 * state: int64 [4] on the device = {next step, current step, address of a pre-generated [T][n] fp32 noise tensor or 0,
 * Philox seed}; the host writes {0, 0, address, seed} before the first step and nothing afterwards.
 * dac_loop_tick (first launch of a step): s = state[0]; t_dev[0] = t_table[s] (the network time of the step, read by
 * dac_time_film); coef_dev[0..8) = coef_table[s][0..8) (coefficients as for dac_sde_step); state[1] = s; state[0] = s+1.
 * dac_sde_step_dev (last launch of a step): dac_sde_step with the coefficients from coef_dev and the noise from row
 * state[1] of the tensor at state[2] (the reference's draws in the parity tests) or, when state[2] == 0, from
 * Philox4x32-10 keyed by (state[3], element / 4, state[1]) inside the kernel (four standard normals per call;
 * statistically, not bitwise, what torch.randn_like draws at SDE:227-231).  n % 4 == 0; out may alias x. */
int dac_loop_tick(int64_t* state, const float* t_table /*[T]*/, const float* coef_table /*[T][8]*/, float* t_dev,
                  float* coef_dev /*[8]*/, dac_stream_t stream);
int dac_sde_step_dev(int mode, const float* x, const float* mu, const float* net, float* out, int64_t n,
                     const float* coef_dev, const int64_t* state, dac_stream_t stream);
/* SDE:374-375  out = x + eps * max_sigma */
int dac_noise_state(const float* x, const float* eps, float* out, int64_t n, float max_sigma, dac_stream_t stream);

/* ------------------------------------------------------------------ UNet stem input  ARCH:123-127
 * xt, cond: [B,3,H,W] fp32 NCHW.  out: [B,Hp,Wp,64] bf16 NHWC where Hp,Wp = H,W reflect-padded up to a
 * multiple of 16 and channel (kx*8 + c), kx in 0..6, c in 0..5, holds cat[xt-cond, cond][c] at column x+kx-3
 * (zero outside the padded image; channels 6,7 of each group and 56..63 are zero).  This turns the 7x7
 * init_conv (ARCH:36,129) into a 7-tap vertical conv over 64 channels for the tensor-core kernel.
 * pair != 0 (Wp even): one packed row per PAIR of horizontally adjacent pixels instead - out [B,Hp,Wp/2,64], group kx'
 * (0..7) holds source column 2 j + kx' - 3; with weights packed as 128 rows (pixel 2j: kx = kx', pixel 2j+1: kx = kx' - 1)
 * init_conv is a 7-tap vertical conv with N = 128 on half as many GEMM rows, its [B,Hp,Wp/2,128] output being the
 * [B,Hp,Wp,64] tensor. */
int dac_unet_stem_input(const float* xt, const float* cond, void* out, int B, int H, int W, int Hp, int Wp,
                        int pair, dac_stream_t stream);

/* ------------------------------------------------------------------ implicit-GEMM conv / linear (tcgen05)
 * One descriptor covers: 3x3 / 1x1 / 7x1 stride-1 convs, the 4x4 stride-2 Downsample (MU:107-108), the
 * nearest-2x-upsample + 3x3 conv (MU:100-104) folded into four 2x2 parity convs, nn.Linear on token tensors,
 * and a virtual channel concat of two sources (ARCH:158,161,167), with the following fused epilogues. */
enum {
  DAC_EPI_PLAIN = 0,  /* out = act(film(acc + bias)) + res                                   MU:115-153 */
  DAC_EPI_GEGLU = 1,  /* tile cols [0,bn/2) value, [bn/2,bn) gate: out = v * gelu(g)        ATT:37-44 */
  DAC_EPI_LN = 2,     /* out = LN_c(acc + bias) * g + res   (single N tile)                 MU:77-86,166-168 */
  DAC_EPI_QKV = 3,    /* N-tile 0: per-32-col softmax * 32^-0.5 (q) -> out NHWC [.,128]; N-tiles 1,2: raw k, v
                         written PLANAR to out_planar [B][256][H][W] (pixel-contiguous rows, the layout the
                         context reduction consumes without a transpose).  cout 128 = q only. MU:170-177 */
  DAC_EPI_KVCTX = 4   /* cout 256 = k | v of LinearAttention, NEVER written to memory: the epilogue forms
                         P = exp(k - kv_shift[d]) and accumulates C[h][d][e] += sum_pixels P[d] v[e] and
                         S[h][d] += sum_pixels P[d] on the warp tensor cores, then STORES them as partial records
                         {C[32][32], m[32] = 0, S[32]} in ctx_acc [B][4][ctx_slots][1088] - the partial format of
                         dac_linattn_fold with nchunks = ctx_slots; slot = 2 * (CTA - first CTA of the image) +
                         epilogue group, so the fold adds them in a fixed order: bit-reproducible, no atomics (see
                         dac_linattn_ctx_slots).  kv_shift is a data-independent upper bound of k
                         (|k_d| <= ||W_k[d]|| sqrt(C) after the gain-free PreNorm), times log2(e)  MU:170-177 */
};
enum { DAC_ACT_NONE = 0, DAC_ACT_SILU = 1, DAC_ACT_GELU = 2 };

typedef struct dac_conv_desc {
  /* input sources (virtual concat along channels): NHWC bf16, c channels used out of pixel pitch ld */
  const void* src0; int32_t c0; int32_t ld0;
  const void* src1; int32_t c1; int32_t ld1;
  int32_t B, H, W;              /* input spatial dims */
  int32_t OH, OW;               /* conv output grid (per parity group) */
  int32_t stride;               /* 1 or 2: input coord = out*stride + tap offset */
  int32_t ngroups;              /* 1, or 4 parity groups for the folded upsample conv */
  int32_t ntaps;                /* taps per group (<=16) */
  /* Taps are organised in "column groups": ncols groups per parity group, each made of ndy vertically adjacent
   * taps (dx, dy0 .. dy0+ndy-1).  One TMA load of (tile_h + ndy - 1) rows serves the ndy taps of a column group:
   * tap i is the same shared-memory tile read through a descriptor shifted by i*tile_w rows.  ndy == 1 gives one
   * load per tap.  col_tap[g][j*ndy+i] = index (< ntaps) of that tap's weight slab within parity group g. */
  int32_t ndy, ncols;
  int8_t col_dx[4][16];
  int8_t col_dy0[4][16];
  int8_t col_tap[4][16];
  int32_t out_scale;            /* output position = out*out_scale + out_off[g] */
  int8_t out_oy[4], out_ox[4];
  /* weights: bf16 [Z][cout_pad][c0+c1], Z = ngroups*ntaps (* B if per_image_w); cout_pad multiple of block_n */
  const void* weight; int32_t cout; int32_t cout_pad; int32_t per_image_w;
  int32_t block_n;              /* UMMA N: 16..256, multiple of 16 */
  /* Optional fused 1x1 skip convolution (ResBlock.res_conv, MU:141,153): out = act(conv(src)) + W_r . rsrc, the
   * second product accumulated in its own TMEM columns by extra K steps over rsrc0|rsrc1 (virtual concat).
   * rweight: bf16 [cout][rc0+rc1].  Needs a single N tile with 2*block_n <= 256 and the PLAIN epilogue. */
  const void* rsrc0; int32_t rc0; int32_t rld0;
  const void* rsrc1; int32_t rc1; int32_t rld1;
  const void* rweight;
  int32_t tile_h, tile_w;       /* tile_h*tile_w == 128 */
  /* epilogue */
  int32_t epi, act;
  const float* bias;            /* [cout] or NULL */
  const float* bias_img;        /* [B][cout] or NULL (per-image bias, e.g. constant cross-attention term) */
  const float* film; int32_t film_ld, film_off;   /* scale at film[b*ld+off+c], shift at +cout; NULL = none */
  const float* ln_g; float ln_eps;
  const void* res; int32_t res_ld;                /* residual NHWC bf16 at output coordinates, or NULL */
  const void* res2; int32_t res2_ld;              /* optional second residual (nested Residual of MU:27-33 + ATT:261) */
  void* out; int32_t out_ld, out_coff;            /* NHWC bf16 [B, OH*out_scale, OW*out_scale, out_ld] */
  const float* res_f32; int32_t res_f32_ld;       /* fp32 residual stream (ViT blocks keep x in fp32), PLAIN only */
  float* out_f32; int32_t out_f32_ld;             /* fp32 copy of the output (may be given together with out) */
  void* out_planar;                               /* DAC_EPI_QKV only: bf16 [B][256][OH][OW] for k | v */
  /* Channel LayerNorm (MU:77-86,89-97 PreNorm) folded around a 1x1 conv: the PRODUCER of x (PLAIN epilogue, one N
   * tile) writes per-pixel {mean, rstd} of its bf16 output row to stats_out; the QKV conv then runs on the raw x
   * with gain-folded weights W' = W diag(g) and finishes rstd * (acc - mean * colsum(W')) in its epilogue, so the
   * normalised tensor is never materialised. */
  float* stats_out; float stats_eps;              /* [B*OH*OW][2] fp32 */
  const float* ln_stats; const float* ln_colsum;  /* QKV, KVCTX: [B*OH*OW][2], [cout]: folded PreNorm of the input rows */
  float* out_nchw; int32_t out_nchw_c, out_nchw_h, out_nchw_w; /* alt. fp32 NCHW output (final_conv), cropped */
  const float* kv_shift; float* ctx_acc;          /* KVCTX: [128] shift * log2(e); [B][4][ctx_slots][1088] fp32 (zeroed by launch) */
  int32_t halo;                                   /* 3x3 stride-1, tile_w 8: ONE (tile_h+2) x (tile_w+2) load per K chunk
                                                     serves all nine taps (descriptors with a (tile_w+2)*128 B group
                                                     stride); ncols = 1, ndy = 9, col_dx = col_dy0 = -1.  Also the 2x2
                                                     taps of each parity group of the folded upsample conv (ntaps = ndy
                                                     = 4, tap i = (i / 2, i % 2) into a (tile_h+1) x (tile_w+1) box whose
                                                     origin is the group's col_dx / col_dy0) */
  int32_t ctx_slots;                              /* KVCTX: partial records per (image, head) in ctx_acc,
                                                     >= dac_linattn_ctx_slots(B, tiles per image, 2) */
  int32_t pair;                                   /* pixel-pair mode of a 3x3 stride-1 conv with 64 output channels (the
                                                     layers an N = 64 MMA caps at 66.6 % of the tensor peak): every tensor
                                                     is passed as its [B, H, W/2, 2C] view (W even; c0 / c1 / cout / block_n
                                                     = twice the real counts, 128), halo loads, weight = bf16
                                                     [3 ky][192 = kx 2,1,0 x 64 cout][cin], FiLM vectors of 64 entries;
                                                     the centre taps run as N = 128 MMAs.  PLAIN epilogue only.
                                                     pair = 2: the weight tensor holds 9 blocks - the layout above followed
                                                     by the two per-rank layouts of the CTA-pair build,
                                                     [3 + 3 rank + ky][192 = E | O | S0 | S2][cin] with, for rank 0 / 1,
                                                     E = W(kx=1) / W(0), O = W(2) / W(1), S0 = W(0)[32 rank, +32),
                                                     S2 = W(2)[32 rank, +32) */
} dac_conv_desc;

/* CTA-pair mode (chosen by dac_conv_create, nothing to request): layers whose weights are streamed, pixel-pair layers
 * (pair = 2) and layers with a fused skip conv run as 2-CTA clusters issuing tcgen05.mma.cta_group::2 (M = 256; every CTA
 * loads its own activation tile and half of the weight rows) when the M-tile count is even; results are bit-identical to the
 * 1-CTA build.  Environment: DAC_CTA2=0 disables it, DAC_CTA2_RES=0 keeps layers with resident non-pair weights on the 1-CTA build. */
typedef struct dac_conv_plan* dac_conv_t;
int dac_conv_create(const dac_conv_desc* desc, dac_conv_t* plan);
int dac_conv_launch(dac_conv_t plan, dac_stream_t stream);
void dac_conv_destroy(dac_conv_t plan);
/* Tiles, CTAs and dynamic shared memory chosen for a plan (for logs and roofline accounting). */
int dac_conv_info(dac_conv_t plan, int32_t* tiles, int32_t* ctas, int32_t* smem_bytes, int32_t* stages);

/* ------------------------------------------------------------------ norms
 * LayerNorm over the last dim of a [rows, c] bf16 matrix (row pitch ld_in / ld_out), fp32 statistics,
 * biased variance.  w/b may be NULL (gain 1 / no bias).  Serves channel LayerNorm (MU:77-86, gain only, NHWC
 * rows = pixels) and nn.LayerNorm of the transformer blocks (ATT:203-205) and ViT (TR:22-28). */
int dac_layernorm_rows(const void* in, int32_t ld_in, void* out, int32_t ld_out, int64_t rows, int32_t c,
                       const float* w, const float* b, float eps, dac_stream_t stream);
/* Same plus a bf16 residual row: out = LayerNorm(in) * w + b + res (res may be NULL).  The `to_out` LayerNorm and the Residual
 * wrapper of a LinearAttention wider than 256 channels (MU:27-33,168,185), whose rows exceed the fused LN epilogue. */
int dac_layernorm_rows_res(const void* in, int32_t ld_in, const void* res, int32_t ld_res, void* out, int32_t ld_out,
                           int64_t rows, int32_t c, const float* w, const float* b, float eps, dac_stream_t stream);
/* GroupNorm(32 groups, eps) over NHWC bf16 [B, hw, c] (ATT:76-77,251).  stats: workspace [B][16][groups][2] fp32 (per-slab
 * partial sums, added in a fixed order: no atomics, bit-reproducible). */
/* Same, fp32 input rows (the ViT residual stream), bf16 output. */
int dac_layernorm_rows_f32(const float* in, int32_t ld_in, void* out, int32_t ld_out, int64_t rows, int32_t c,
                           const float* w, const float* b, float eps, dac_stream_t stream);
int dac_groupnorm_nhwc(const void* in, void* out, int32_t B, int32_t hw, int32_t c, int32_t groups,
                       const float* w, const float* b, float eps, float* stats, dac_stream_t stream);
/* PreNorm + GroupNorm of a SpatialTransformer level in two launches: normed = LayerNorm_c(in) * pre_g (MU:77-97; kept: it is
 * the residual of proj_out, ATT:261) and out = GroupNorm(normed) (ATT:76-77,251).  The first kernel normalises the rows AND
 * accumulates the per-slab group sums of the bf16 values it stores; stats as for dac_groupnorm_nhwc. */
int dac_prenorm_groupnorm_nhwc(const void* in, void* normed, void* out, int32_t B, int32_t hw, int32_t c, int32_t groups,
                               const float* pre_g, float pre_eps, const float* w, const float* b, float eps, float* stats,
                               dac_stream_t stream);

/* ------------------------------------------------------------------ conditioning vectors
 * time_mlp + text_mlp/prompt/prompt_mlp (ARCH:51-62,132-137) -> silu(t_emb) [B,256] fp32, then every ResBlock
 * `mlp` Linear at once (MU:135-137,145-148): film[b, f] = W_all[f,:] . silu(t_emb[b]) + b_all[f].
 * weights: fp32 row-major in the reference state-dict layout. */
typedef struct dac_embed_weights {
  const float *time_w1, *time_b1, *time_w2, *time_b2;       /* [256,64],[256],[256,256],[256] */
  const float *text_w1, *text_b1, *text_w2, *text_b2;       /* [256,ctx],[256],[256,256],[256] (NULL: no prompt) */
  const float *prompt, *prompt_w, *prompt_b;                /* [256],[256,256],[256] */
  const float *film_w, *film_b;                             /* [F,256],[F] */
  int32_t nf, time_dim, ctx_dim, F;
} dac_embed_weights;
/* Per restoration (step-invariant): prompt_emb[b] = prompt_mlp(softmax(text_mlp(text_ctx[b])) * prompt)  ARCH:134-136 */
int dac_prompt_embed(const dac_embed_weights* w, const float* text_ctx /*[B,ctx]*/, int32_t B,
                     float* prompt_emb /*[B,time_dim]*/, dac_stream_t stream);
/* Per step: silu(time_mlp(t) + prompt_emb[b]) -> temb_scratch, then the stacked ResBlock mlp linears -> film. */
int dac_time_film(const dac_embed_weights* w, const float* time /*device scalar: one CUDA graph serves every step*/,
                  const float* prompt_emb /*[B,time_dim] or NULL*/, int32_t B,
                  float* temb_scratch /*[B,time_dim]*/, float* film /*[B,F]*/, dac_stream_t stream);
/* out[b, :] = W2 (W1 x[b]) + b2  -- the exact value of cross-attention over a 1-token context
 * (ATT:152-193 with len(context)=1: softmax of one logit == 1).  W1 [mid,in], W2 [out,mid] fp32. */
int dac_two_linear(const float* x, int32_t B, int32_t in, const float* w1, int32_t mid, const float* w2,
                   const float* b2, int32_t out, float* y, dac_stream_t stream);

/* ------------------------------------------------------------------ LinearAttention (MU:157-185)
 * kv: planar bf16 [B][256][hw] (channels 0..127 k, 128..255 v) from the QKV epilogue.  Pass 1 reduces, per (image, head),
 * the k-softmax over all pixels and ctx[d,e] = sum_n softmax_n(k)[d,n] v[e,n] / hw into partials; pass 2
 * merges them and folds ctx into the to_out weight: weff[b][c][h*32+d] = sum_e Wout[c][h*32+e] ctx[b,h,d,e]
 * (bf16, the per-image weight of the following 1x1 conv plan). */
int dac_linattn_context(const void* kv, int32_t B, int32_t hw, int32_t nchunks, float* partial /*[B,4,nchunks,32*34]*/,
                        dac_stream_t stream);
int dac_linattn_fold(const float* partial, int32_t B, int32_t hw, int32_t nchunks, const float* w_out /*[C,128] fp32*/,
                     int32_t C, int32_t c_pad, void* weff /*[B][c_pad][128] bf16*/, dac_stream_t stream);

/* The fold for dac_linattn_kv's in-kernel-PreNorm mode: partial = [B][4][nslots][2080] records {G[32][64], S[32]};
 * m_fold = fp32 [4][C][64], M_h[c'][c] = sum_e W_out[c'][h*32+e] W_vc[h*32+e][c] (W_vc = the bf16 gain-folded, row-centred
 * value rows of to_qkv; rows of M_h centred again in fp32); weff[b][c'][h*32+d] = sum_c G[b,h,d,c] M_h[c'][c] / (S[b,h,d] hw) (MU:170-184 with the context never formed). */
int dac_linattn_fold_g(const float* partial, int32_t B, int32_t hw, int32_t nslots, const float* m_fold /*[4,C,64] fp32*/,
                       int32_t C, int32_t c_pad, void* weff /*[B][c_pad][128] bf16*/, dac_stream_t stream);

/* Partial records per (image, head) that the context-reducing kernels need: an image's tiles are contiguous in the tile
 * order and CTA b owns tiles [tiles*b/grid, tiles*(b+1)/grid), grid = min(tiles, SM count); every CTA that touches an
 * image stores `groups` partials for it (2 for DAC_EPI_KVCTX - one per epilogue group - 1 for dac_linattn_kv).  Returns
 * groups * (largest number of CTAs any image spans). */
int32_t dac_linattn_ctx_slots(int32_t B, int32_t tiles_per_image, int32_t groups);

/* Key/value side of LinearAttention entirely on tcgen05 (MU:170-177): per 128-pixel tile k | v = wkv . xn, P = exp(k - c_d)
 * (kv_shift as for DAC_EPI_KVCTX), then C += P^T V and S += P^T 1 as a second tensor-core GEMM with MN-major operands,
 * accumulated in tensor memory over the CTA's tiles of an image and stored as one partial record per (CTA, image) into
 * ctx_acc [B][4][ctx_slots][1088] fp32 (zeroed by the launch; ctx_slots >= dac_linattn_ctx_slots(B, hw / 128, 1);
 * merged in slot order by dac_linattn_fold with nchunks = ctx_slots).  xn: bf16 [B*hw, C] (C = 64 or 128, hw % 128 == 0);
 * wkv: bf16 [256][C], rows packed per head pair g as k_2g k_2g+1 v_2g v_2g+1.  Replaces DAC_EPI_KVCTX where it fits.
 * Folded PreNorm (MU:89-97): with ln_stats ([B*hw][2] fp32 {mean, rstd} per pixel, written by the producing layer's
 * stats_out) and ln_colsum ([256] fp32 row sums of the bf16 wkv rows, same packed order) `xn` is the RAW input and the
 * epilogue finishes the normalisation: W' LN(x) = rstd (W' x - mean colsum(W')); both NULL: xn is already normalised.
 * In-kernel PreNorm (prenorm != 0; C = 64, ln_stats NULL): `xn` is the RAW input and the gain-free channel LayerNorm
 * (eps prenorm_eps, MU:77-86) is folded into the GEMMs: wkv holds ONLY the key rows, ROW-CENTRED (W_k - rowmean(W_k), so
 * that W_k LN(x) = rstd (wkv . x)), bf16 [128][64] in head order; four warps compute each pixel's rstd from the landed
 * tile - no LayerNorm launch, no normalised tensor.  The values are never formed: the kernel accumulates
 * G'[(h,d)][c] = sum_px P[px][(h,d)] rstd[px] x[px][c] and S = sum_px P (context = G' W_vc^T / S with the row-centred
 * W_vc); ctx_acc is [B][4][ctx_slots][2080] fp32 records {G'[32][64], S[32]}, to be merged by dac_linattn_fold_g (NOT
 * dac_linattn_fold). */
typedef struct dac_kv_plan* dac_kv_t;
int dac_linattn_kv_create(const void* xn, const void* wkv, const float* kv_shift, float* ctx_acc, int32_t ctx_slots,
                          const float* ln_stats, const float* ln_colsum, int32_t B, int32_t hw, int32_t C,
                          int32_t prenorm, float prenorm_eps, dac_kv_t* plan);
int dac_linattn_kv_launch(dac_kv_t plan, dac_stream_t stream);
void dac_linattn_kv_destroy(dac_kv_t plan);

/* Query side of LinearAttention as ONE chained-GEMM kernel (MU:170-185): per 128-pixel tile
 *   q = softmax_head-channels(wq . xn) * 32^-0.5   (kept in shared memory, bf16)
 *   out = LayerNorm_c(weff[b] . q + bias) * ln_g + res
 * xn, res, out: bf16 [B*hw, C] (C = 64 or 128, hw % 128 == 0); wq: bf16 [128][C] (gain-folded rows of to_qkv);
 * weff: bf16 [B][c_pad][128] from dac_linattn_fold.  Replaces the to_q (DAC_EPI_QKV) + to_out (DAC_EPI_LN) pair.
 * ln_stats / ln_colsum ([B*hw][2], [128]; or both NULL): folded PreNorm as for dac_linattn_kv_create - xn is then the
 * raw input (normally the same tensor as res).
 * In-kernel PreNorm (prenorm != 0; C = 64, ln_stats NULL, res == xn): `xn` is the RAW input and wq holds the ROW-CENTRED
 * rows W_q - rowmean(W_q), so that W_q LN(x) = rstd (wq . x): every 128 x 64 tile is read ONCE, is the A operand of the
 * first GEMM as it is, gets its per-pixel rstd (eps prenorm_eps, MU:77-86) from four statistics warps, and is reused as
 * the residual - no LayerNorm launch, no normalised tensor, half the input traffic.  q_shift (fp32 [128] on the device,
 * or NULL): data-independent bounds c_d * log2(e) >= |q_d| log2(e) (c_d = ||wq[d]|| sqrt(C)); the softmax over a head's
 * channels then uses max_d c_d as its shift instead of the row maximum (needs c_d <= 40). */
typedef struct dac_qout_plan* dac_qout_t;
int dac_linattn_qout_create(const void* xn, const void* wq, const void* weff, int32_t c_pad, const void* res,
                            void* out, const float* bias, const float* ln_g, float ln_eps, const float* ln_stats,
                            const float* ln_colsum, int32_t B, int32_t hw, int32_t C, int32_t prenorm,
                            float prenorm_eps, const float* q_shift, dac_qout_t* plan);
int dac_linattn_qout_launch(dac_qout_t plan, dac_stream_t stream);
void dac_linattn_qout_destroy(dac_qout_t plan);

/* ------------------------------------------------------------------ softmax attention
 * qkv: [B, n, 3*heads*d] bf16 packed (q | k | v along channels), out [B, n, heads*d] bf16.
 * d = 32 (UNet self-attention, ATT:178-192) or 64 (ViT, TR:219-230); scale = d^-0.5. */
int dac_attention(const void* qkv, void* out, int32_t B, int32_t n, int32_t heads, int32_t d, dac_stream_t stream);

/* Causal variant for the CLIP text tower (OC/model.py:237-249; mask = build_attention_mask, OC/transformer.py:
 * upper triangle -inf): query i attends keys 0..i.  d = 64 only (text widths 512 / 768 with 8 / 12 heads). */
int dac_attention_causal(const void* qkv, void* out, int32_t B, int32_t n, int32_t heads, int32_t d,
                         dac_stream_t stream);

/* ------------------------------------------------------------------ DA-CLIP encoder helpers (TR:507-555)
 * patchify: image [B,3,S,S] fp32 NCHW -> [B*g*g, 3*p*p] bf16 rows (k = c*p*p + py*p + px, conv1 weight order) at a
 * row pitch of ld_out >= 3*p*p elements (the GEMM wants K % 64 == 0: ViT-L/14 pads 588 -> 640; pad columns are left
 * untouched, the caller zeroes them once). */
int dac_vit_patchify(const float* image, void* out, int32_t B, int32_t S, int32_t p, int32_t ld_out,
                     dac_stream_t stream);
/* tokens[b,0,:] = cls + pos[0]; tokens[b,1+i,:] = patch[b,i,:] + pos[1+i]; then ln_pre -> out fp32 [B,L,w]. */
int dac_vit_embed(const void* patch_emb, const float* cls, const float* pos, const float* ln_w, const float* ln_b,
                  void* out, int32_t B, int32_t L, int32_t w, float eps, dac_stream_t stream);
/* pooled[b] = ln_post(x[b,0,:]) @ proj  -> fp32 [B, e];  x: fp32 [B,L,w] */
int dac_vit_pool(const void* x, int32_t B, int32_t L, int32_t w, const float* ln_w, const float* ln_b, float eps,
                 const float* proj /*[w,e]*/, int32_t e, float* out, dac_stream_t stream);
/* Text tower entry and exit (CLIP.encode_text, OC/model.py:237-249):
 * embed: out[b,t,:] = token_embedding[text[b,t]] + pos[t] (fp32 [B,L,w]); eot[b] = argmax_t text[b,t] (first maximum).
 *        text: int64 [B,L] token ids in [0, vocab) (the binding validates them; the reference raises on others).
 * pool:  out[b] = ln_final(x[b, eot[b], :]) @ text_projection -> fp32 [B, e]. */
int dac_text_embed(const int64_t* text, const float* token_embedding /*[vocab,w]*/, const float* pos /*[L,w]*/,
                   float* out, int32_t* eot /*[B]*/, int32_t B, int32_t L, int32_t w, int32_t vocab,
                   dac_stream_t stream);
int dac_text_pool(const void* x, const int32_t* eot, int32_t B, int32_t L, int32_t w, const float* ln_w,
                  const float* ln_b, float eps, const float* proj /*[w,e]*/, int32_t e, float* out,
                  dac_stream_t stream);
/* Degradation-type argmax (da-clip/src/evaluate_daclip.py:46-47,79-81): argmax_j 100*cos(degra[b], text[j]). */
int dac_degradation_argmax(const float* degra, const float* text, int32_t B, int32_t e, int32_t classes,
                           float* logits /*[B,classes] or NULL*/, int64_t* argmax, dac_stream_t stream);

/* ------------------------------------------------------------------ image pre / post-processing
 * clip_transform (universal-image-restoration/data/util.py:87-93): float RGB HWC image in [0,1] -> (uint8)(v*255) ->
 * Pillow BICUBIC resize (antialiased, 8-bit fixed point: `bounds` [out][2] = {first tap, tap count}, `kk`
 * [out][ksize] = coefficients * 2^22, computed on the host as Pillow's precompute_coeffs / normalize_coeffs_8bpc do)
 * -> CenterCrop -> ToTensor -> Normalize.  Pass 1 resamples rows [y_first, y_first + rows) horizontally into the
 * uint8 buffer `mid` [rows][Wout][3]; pass 2 resamples vertically, crops res x res at (top, left) and writes
 * (x / 255 - mean) / std as fp32 [3][res][res].  Bit-exact against Pillow.  mean3 / std3 are HOST pointers. */
int dac_clip_resample_h(const float* img, int32_t H, int32_t W, void* mid, int32_t Wout, const int32_t* bounds,
                        const int32_t* kk, int32_t ksize, int32_t y_first, int32_t rows, dac_stream_t stream);
int dac_clip_resample_v_norm(const void* mid, int32_t Wout, int32_t y_first, const int32_t* bounds, const int32_t* kk,
                             int32_t ksize, int32_t top, int32_t left, int32_t res, const float* mean3,
                             const float* std3, float* out, dac_stream_t stream);
/* tensor2img (universal-image-restoration/utils/img_utils.py:136-163): x fp32 [B][C][H][W], C = 3 (RGB) or 1 ->
 * uint8 [B][H][W][C], channels reversed (BGR): round_half_even((clamp(x, lo, hi) - lo) / (hi - lo) * 255). */
int dac_tensor2img(const float* x, void* out, int32_t B, int32_t C, int32_t H, int32_t W, float lo, float hi,
                   dac_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DAC_B200_H */
