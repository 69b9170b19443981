// Per-pixel channel-LayerNorm statistics of one 64-channel row of a 128B-swizzled shared-memory tile, for the kernels that
// fold the PreNorm of LinearAttention (module_util.py:77-97) into their GEMMs:
//   W LN(x) = rstd * (W_c x),  W_c = W - rowmean(W)   (the row sums of W_c vanish, so the mean term drops out and the
// RAW bf16 tile is the GEMM operand; only rstd - one scalar per accumulator row - is left for the epilogue).
#pragma once
#include "ptx.cuh"

namespace dac {

// Row `row` of the tile at shared address `tile`: the 128 bytes at row * 128 with their 16-byte pieces permuted (piece j at
// j ^ (row & 7)).  The moments do not care about the order; step j touches piece j ^ (row & 7), so the eight rows of a
// quarter-warp hit eight different bank groups.  Returns mean and the biased variance E[x^2] - mean^2 (fp32, >= 0).
__device__ __forceinline__ void row_moments64(uint32_t tile, int row, float& mean, float& var) {
  const uint32_t base = tile + row * 128;
  uint64_t s2[2] = {0ull, 0ull}, q2[2] = {0ull, 0ull};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint32_t u[4];
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3])
                 : "r"(base + 16 * (j ^ (row & 7))));
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint64_t x = pack_f32x2(__uint_as_float(u[q] << 16), __uint_as_float(u[q] & 0xffff0000u));
      s2[q & 1] = add_f32x2(s2[q & 1], x);
      q2[q & 1] = fma_f32x2(x, x, q2[q & 1]);
    }
  }
  float a, b;
  unpack_f32x2(add_f32x2(s2[0], s2[1]), a, b);
  mean = (a + b) * (1.0f / 64.0f);
  unpack_f32x2(add_f32x2(q2[0], q2[1]), a, b);
  var = fmaxf(fmaf(-mean, mean, (a + b) * (1.0f / 64.0f)), 0.f);
}

}  // namespace dac
