// Geometry helpers shared by the elementwise stages: odometry (de)normalisation
// (transfuser_model_v2.py:480-500) and the bilinear corners of grid_sample
// (modules/blocks.py:101-122).  Intrinsics keep the reference's operation order (no FMA
// contraction) so fp32 results track torch's.
#pragma once
#include "kernels.h"

namespace ddh {

__device__ __forceinline__ float norm_x(float x) {   // 2*(x+1.2)/56.9 - 1
  return __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fadd_rn(x, 1.2f)), 56.9f), 1.0f);
}
__device__ __forceinline__ float norm_y(float y) {   // 2*(y+20)/46 - 1
  return __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fadd_rn(y, 20.0f)), 46.0f), 1.0f);
}
__device__ __forceinline__ float denorm_x(float v) { // (v+1)/2*56.9 - 1.2
  return __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn(v, 1.0f), 2.0f), 56.9f), 1.2f);
}
__device__ __forceinline__ float denorm_y(float v) { // (v+1)/2*46 - 20
  return __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn(v, 1.0f), 2.0f), 46.0f), 20.0f);
}

struct Corners {
  int pix[4];
  float w[4];
};
__device__ __forceinline__ Corners corners_of(float px, float py, int H, int W, OdoConsts oc) {
  Corners c;
  const float gx = __fdiv_rn(py, oc.lidar_max_x);
  const float gy = __fdiv_rn(px, oc.lidar_max_y);
  const float ix = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.0f), (float)W), 1.0f), 2.0f);
  const float iy = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.0f), (float)H), 1.0f), 2.0f);
#pragma unroll
  for (int k = 0; k < 4; ++k) { c.pix[k] = -1; c.w[k] = 0.f; }
  if (!(ix > -2.0f && ix < (float)(W + 1) && iy > -2.0f && iy < (float)(H + 1))) return c;
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const int x0 = (int)fx0, y0 = (int)fy0;
  const float wx1 = __fsub_rn(ix, fx0), wx0 = __fsub_rn(__fadd_rn(fx0, 1.0f), ix);
  const float wy1 = __fsub_rn(iy, fy0), wy0 = __fsub_rn(__fadd_rn(fy0, 1.0f), iy);
  const int xs[4] = {x0, x0 + 1, x0, x0 + 1};
  const int ys[4] = {y0, y0, y0 + 1, y0 + 1};
  const float ws[4] = {__fmul_rn(wx0, wy0), __fmul_rn(wx1, wy0), __fmul_rn(wx0, wy1),
                       __fmul_rn(wx1, wy1)};  // nw, ne, sw, se
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (xs[k] >= 0 && xs[k] < W && ys[k] >= 0 && ys[k] < H) {
      c.pix[k] = ys[k] * W + xs[k];
      c.w[k] = ws[k];
    }
  }
  return c;
}

}  // namespace ddh
