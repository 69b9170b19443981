// Host-side pre/post-processing of the reference's test loop, moved onto the GPU:
//   clip_transform  (universal-image-restoration/data/util.py:87-93): float RGB HWC -> uint8 -> PIL bicubic
//                   (antialiased) resize of the short side to 224 -> centre crop -> ToTensor -> Normalize;
//   tensor2img      (universal-image-restoration/utils/img_utils.py:136-163): clamp -> [0,1] -> *255 -> round -> uint8,
//                   RGB CHW -> BGR HWC.
// Integer / byte work, bit-exact against Pillow's 8-bit resampler: the per-output-pixel tap ranges and the 22-bit
// fixed-point coefficients are computed on the host exactly as ImagingResample does (da-clip_b200/imageio.py) and
// the two passes below accumulate in int32 like the C code, with the uint8 intermediate between them.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dac_b200.h"
#include "common.h"

namespace dac {

constexpr int kPrecisionBits = 32 - 8 - 2;   // Pillow: PRECISION_BITS

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;                     // arithmetic shift, as the C code
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// Horizontal pass.  Source = the float image (quantised on the fly: (uint8)(v * 255), util.py:88); one thread per
// (row, output column), three channels.
__global__ void __launch_bounds__(256) resample_h_kernel(const float* __restrict__ img, int H, int W,
                                                         uint8_t* __restrict__ mid, int Wout,
                                                         const int32_t* __restrict__ bounds,
                                                         const int32_t* __restrict__ kk, int ksize, int y_first,
                                                         int rows) {
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (xx >= Wout || r >= rows) return;
  const int y = y_first + r;
  const int xmin = bounds[2 * xx], xmax = bounds[2 * xx + 1];
  const int32_t* k = kk + static_cast<int64_t>(xx) * ksize;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  const float* row = img + (static_cast<int64_t>(y) * W + xmin) * 3;
  for (int x = 0; x < xmax; ++x) {
    const int w = __ldg(k + x);
    // astype(np.uint8) of float32 * 255: truncation toward zero, modulo 256
    s0 += static_cast<int>(static_cast<uint8_t>(static_cast<int>(__fmul_rn(__ldg(row + 3 * x), 255.f)))) * w;
    s1 += static_cast<int>(static_cast<uint8_t>(static_cast<int>(__fmul_rn(__ldg(row + 3 * x + 1), 255.f)))) * w;
    s2 += static_cast<int>(static_cast<uint8_t>(static_cast<int>(__fmul_rn(__ldg(row + 3 * x + 2), 255.f)))) * w;
  }
  uint8_t* o = mid + (static_cast<int64_t>(r) * Wout + xx) * 3;
  o[0] = clip8(s0);
  o[1] = clip8(s1);
  o[2] = clip8(s2);
}

// Vertical pass fused with CenterCrop + ToTensor + Normalize: one thread per output pixel of the crop.
__global__ void __launch_bounds__(256) resample_v_norm_kernel(const uint8_t* __restrict__ mid, int Wout, int y_first,
                                                              const int32_t* __restrict__ bounds,
                                                              const int32_t* __restrict__ kk, int ksize, int top,
                                                              int left, int res, float m0, float m1, float m2,
                                                              float d0, float d1, float d2, float* __restrict__ out) {
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  const int yy = blockIdx.y;
  if (xx >= res || yy >= res) return;
  const int oy = yy + top, ox = xx + left;
  const int ymin = bounds[2 * oy], ymax = bounds[2 * oy + 1];
  const int32_t* k = kk + static_cast<int64_t>(oy) * ksize;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  const uint8_t* col = mid + (static_cast<int64_t>(ymin - y_first) * Wout + ox) * 3;
  for (int y = 0; y < ymax; ++y) {
    const int w = __ldg(k + y);
    const uint8_t* px = col + static_cast<int64_t>(y) * Wout * 3;
    s0 += static_cast<int>(px[0]) * w;
    s1 += static_cast<int>(px[1]) * w;
    s2 += static_cast<int>(px[2]) * w;
  }
  const int64_t plane = static_cast<int64_t>(res) * res;
  float* o = out + static_cast<int64_t>(yy) * res + xx;
  // ToTensor: uint8 / 255 (fp32); Normalize: (x - mean) / std
  o[0] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(clip8(s0)), 255.f), m0), d0);
  o[plane] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(clip8(s1)), 255.f), m1), d1);
  o[2 * plane] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(clip8(s2)), 255.f), m2), d2);
}

// tensor2img: fp32 [B][C][H][W] (C = 3 RGB or 1) -> uint8 [B][H][W][C] with channels reversed (BGR) for C = 3.
// One thread per pixel: coalesced planar reads, 3-byte interleaved write.
__global__ void __launch_bounds__(256) tensor2img_kernel(const float* __restrict__ x, uint8_t* __restrict__ out,
                                                         int C, int64_t hw, float lo, float hi) {
  const int b = blockIdx.y;
  const float range = __fsub_rn(hi, lo);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < hw;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    for (int c = 0; c < C; ++c) {
      float v = __ldg(x + (static_cast<int64_t>(b) * C + c) * hw + i);
      v = fminf(fmaxf(v, lo), hi);                                        // clamp_(*min_max)  (NaN -> min here)
      v = __fdiv_rn(__fsub_rn(v, lo), range);                             // to [0, 1]
      const float q = rintf(__fmul_rn(v, 255.0f));                        // numpy round(): half to even
      out[(static_cast<int64_t>(b) * hw + i) * C + (C - 1 - c)] = static_cast<uint8_t>(static_cast<int>(q));
    }
  }
}

}  // namespace dac

using namespace dac;

extern "C" int dac_clip_resample_h(const float* img, int32_t H, int32_t W, void* mid, int32_t Wout,
                                   const int32_t* bounds, const int32_t* kk, int32_t ksize, int32_t y_first,
                                   int32_t rows, dac_stream_t stream) {
  if (!img || !mid || !bounds || !kk) return set_error(-1, "dac_clip_resample_h: null argument");
  if (H <= 0 || W <= 0 || Wout <= 0 || ksize <= 0 || y_first < 0 || rows <= 0 || y_first + rows > H || rows > 65535)
    return set_error(-2, "dac_clip_resample_h: bad sizes");
  resample_h_kernel<<<dim3(static_cast<unsigned>(ceil_div(Wout, 256)), rows), 256, 0,
                      static_cast<cudaStream_t>(stream)>>>(img, H, W, static_cast<uint8_t*>(mid), Wout, bounds, kk,
                                                           ksize, y_first, rows);
  return check_launch("resample_h_kernel");
}

extern "C" int dac_clip_resample_v_norm(const void* mid, int32_t Wout, int32_t y_first, const int32_t* bounds,
                                        const int32_t* kk, int32_t ksize, int32_t top, int32_t left, int32_t res,
                                        const float* mean3, const float* std3, float* out, dac_stream_t stream) {
  if (!mid || !bounds || !kk || !mean3 || !std3 || !out) return set_error(-1, "dac_clip_resample_v_norm: null argument");
  if (Wout <= 0 || ksize <= 0 || top < 0 || left < 0 || res <= 0 || left + res > Wout || res > 65535)
    return set_error(-2, "dac_clip_resample_v_norm: bad sizes");
  resample_v_norm_kernel<<<dim3(static_cast<unsigned>(ceil_div(res, 256)), res), 256, 0,
                           static_cast<cudaStream_t>(stream)>>>(static_cast<const uint8_t*>(mid), Wout, y_first, bounds,
                                                                kk, ksize, top, left, res, mean3[0], mean3[1],
                                                                mean3[2], std3[0], std3[1], std3[2], out);
  return check_launch("resample_v_norm_kernel");
}

extern "C" int dac_tensor2img(const float* x, void* out, int32_t B, int32_t C, int32_t H, int32_t W, float lo,
                              float hi, dac_stream_t stream) {
  if (!x || !out) return set_error(-1, "dac_tensor2img: null argument");
  if (B <= 0 || (C != 1 && C != 3) || H <= 0 || W <= 0 || !(hi > lo) || B > 65535)
    return set_error(-2, "dac_tensor2img: need C in {1,3}, positive sizes, max > min");
  const int64_t hw = static_cast<int64_t>(H) * W;
  int64_t blocks = ceil_div(hw, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  tensor2img_kernel<<<dim3(static_cast<unsigned>(blocks), B), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<uint8_t*>(out), C, hw, lo, hi);
  return check_launch("tensor2img_kernel");
}
