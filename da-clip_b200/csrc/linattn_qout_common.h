// Shared between linattn_qout.cu (plan creation, round-1 kernel) and linattn_qout2.cu (in-kernel PreNorm, 16 epilogue warps).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace dac {
struct Qout2Params {
  int tiles, tiles_per_image, c_pad;
  int use_max;             // 1: softmax shift = the row maximum; 0: the data-independent per-head bound q_shift
  float ln_eps;            // eps of the output LayerNorm
  float prenorm_eps;       // eps of the channel LayerNorm folded into GEMM 1
  float q_shift[4];        // per head: max_d c_d * log2(e)
  float bias[64];          // to_out bias (zeros if absent): in the parameter constant bank, i.e. instruction operands
  float ln_g[64];          // gain of the output LayerNorm
};
}  // namespace dac

int dac_qout2_smem_bytes();
int dac_qout2_launch(const CUtensorMap& mapX, const CUtensorMap& mapWq, const CUtensorMap& mapWeff,
                     const CUtensorMap& mapOut, const dac::Qout2Params& kp, int grid, cudaStream_t st);
