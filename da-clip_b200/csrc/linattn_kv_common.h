// Shared between linattn_kv.cu (plan creation, round-1 kernel) and linattn_kv2.cu (in-kernel PreNorm, G = P^T xn instead of the context) and linattn.cu (dac_linattn_fold_g).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace dac {
constexpr int kKvGRec = 32 * 64 + 32;   // {G[32 d][64 c], S[32]} per (image, head, slot): the record of dac_linattn_fold_g
struct Kv2Params {
  int tiles, tiles_per_image;
  float shift_max[4];      // per head: max_d c_d * log2(e)
  float* ctx_acc;          // [B][4][slots][kKvGRec]: one partial record per (CTA, image)
  int slots;
  float ln_eps;            // eps of the channel LayerNorm applied to the raw input rows
  long long* prof;         // DAC_KV2_PROF builds: [22 warps][8] cycle totals of CTA 0
  int dbg;                 // timing experiments only (DAC_KV2_DBG): 1 no exp2, 2 no GEMM 2, 4 no P staging, 8 no row moments
};
}  // namespace dac

int dac_kv2_smem_bytes();
int dac_kv2_launch(const CUtensorMap& mapX, const CUtensorMap& mapW, const dac::Kv2Params& kp, int grid, cudaStream_t st);
