// cross_bev_feature producer: the caller-side stage of
// V2TransfuserModel.forward (transfuser_model_v2.py:121-140) that builds the head's BEV input
//
//   up    = bilinear_8x8->HxW(keyval tokens as a (256, g, g) grid, align_corners=False)     (:121-131)
//   x     = cat([up, bev_map], channel)                                                      (:133-134)
//   out   = LayerNorm(ReLU(Linear(256 + Cb -> 256)(x)))  per pixel                           (:136-137, :96)
//   (the reference then permutes NHWC -> NCHW, :138-140; the head wants NHWC, so it is skipped)
//
// fused into two launches.  Linear and bilinear interpolation commute, so the 256-channel half of
// the Linear is applied to the g*g tokens first (kernel 1: 64 rows instead of 4096 pixels) and the
// projected tokens are interpolated per pixel; the upsampled 256-channel map, the 320-channel
// concat and the permute never exist.  Kernel 2 is one warp per 4 pixels: the Cb x 256 slice of
// the weights sits in shared memory, the Cb map channels of a 32-pixel run are staged coalesced,
// LayerNorm is a warp reduction, the result is written as NHWC bf16 (or fp32) lines.
#include "kernels.h"

namespace ddh {
namespace {

// Pt[b][t][c] = sum_k tok[b][t][k] * w[c][k]  (k < 256) ; WbT[j][c] = w[c][256 + j]
__global__ void __launch_bounds__(256) producer_tokens_kernel(const float* __restrict__ tok,
                                                              const float* __restrict__ w,
                                                              float* __restrict__ Pt, float* __restrict__ WbT,
                                                              int n_tok, int ldw, int cb) {
  __shared__ float ts[8][256];
  __shared__ float ws[32][257];
  const int b = blockIdx.y, t0 = blockIdx.x * 8, c = threadIdx.x;
  for (int i = threadIdx.x; i < 8 * 256; i += 256) {
    const int t = t0 + (i >> 8);
    ts[i >> 8][i & 255] = t < n_tok ? tok[((size_t)b * n_tok + t) * 256 + (i & 255)] : 0.f;
  }
  if (b == 0 && blockIdx.x == 0)
    for (int i = threadIdx.x; i < cb * 256; i += 256) {
      const int cc = i / cb, j = i - cc * cb;       // coalesced along j
      WbT[(size_t)j * 256 + cc] = w[(size_t)cc * ldw + 256 + j];
    }
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int k0 = 0; k0 < 256; k0 += 32) {
    __syncthreads();
    // thread c stages its own weight row piece (128 contiguous bytes), transposed into ws[k][c]
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(w + (size_t)c * ldw + k0) + q);
      ws[4 * q + 0][c] = v.x; ws[4 * q + 1][c] = v.y; ws[4 * q + 2][c] = v.z; ws[4 * q + 3][c] = v.w;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float wv = ws[k][c];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(ts[i][k0 + k], wv, acc[i]);
    }
  }
  for (int i = 0; i < 8; ++i)
    if (t0 + i < n_tok) Pt[((size_t)b * n_tok + t0 + i) * 256 + c] = acc[i];
}

template <typename TO>
__global__ void __launch_bounds__(256) producer_pixels_kernel(
    const float* __restrict__ Pt, const float* __restrict__ WbT, const float* __restrict__ map,
    const float* __restrict__ bias, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
    TO* __restrict__ out, int H, int W, int g, int cb, int n_runs) {
  extern __shared__ __align__(16) float psm[];
  float* wb = psm;                    // [cb][256]
  float* fs = psm + (size_t)cb * 256; // [cb][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < cb * 64; i += 256)
    reinterpret_cast<float4*>(wb)[i] = __ldg(reinterpret_cast<const float4*>(WbT) + i);
  float bv[8], gv[8], ev[8];
  load8(bias, lane, bv);
  load8(ln_g, lane, gv);
  load8(ln_b, lane, ev);
  const int HW = H * W, runs_per_scene = HW / 32;
  const float sy = (float)g / (float)H, sx = (float)g / (float)W;
  for (int run = blockIdx.x; run < n_runs; run += gridDim.x) {
    const int b = run / runs_per_scene, px0 = (run - b * runs_per_scene) * 32;
    __syncthreads();
    for (int i = threadIdx.x; i < cb * 8; i += 256) {   // cb channels x 32 pixels, 128-byte lines
      const int j = i >> 3, q = i & 7;
      reinterpret_cast<float4*>(fs)[i] = __ldg(reinterpret_cast<const float4*>(map + ((size_t)b * cb + j) * HW + px0) + q);
    }
    __syncthreads();
    const int p0 = warp * 4;            // this warp's 4 pixels of the run
    float acc[4][8];
    // ---- bilinear interpolation of the projected tokens (F.interpolate, align_corners=False)
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int px = px0 + p0 + p, y = px / W, x = px - y * W;
      const float fy = fmaxf(((float)y + 0.5f) * sy - 0.5f, 0.f), fx = fmaxf(((float)x + 0.5f) * sx - 0.5f, 0.f);
      const int y0 = (int)fy, x0 = (int)fx;
      const int y1 = min(y0 + 1, g - 1), x1 = min(x0 + 1, g - 1);
      const float ly = fy - (float)y0, lx = fx - (float)x0;
      const float* base = Pt + (size_t)b * g * g * 256;
      float t00[8], t01[8], t10[8], t11[8];
      load8(base + (size_t)(y0 * g + x0) * 256, lane, t00);
      load8(base + (size_t)(y0 * g + x1) * 256, lane, t01);
      load8(base + (size_t)(y1 * g + x0) * 256, lane, t10);
      load8(base + (size_t)(y1 * g + x1) * 256, lane, t11);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        // torch's upsample_bilinear2d order: (1-ly) * ((1-lx) a + lx b) + ly * ((1-lx) c + lx d)
        const float top = (1.f - lx) * t00[i] + lx * t01[i];
        const float bot = (1.f - lx) * t10[i] + lx * t11[i];
        acc[p][i] = bv[i] + ((1.f - ly) * top + ly * bot);
      }
    }
    // ---- the map's cb channels
    for (int j = 0; j < cb; ++j) {
      const float4 w0 = *reinterpret_cast<const float4*>(wb + (size_t)j * 256 + lane * 4);
      const float4 w1 = *reinterpret_cast<const float4*>(wb + (size_t)j * 256 + 128 + lane * 4);
      const float4 f4 = *reinterpret_cast<const float4*>(fs + j * 32 + p0);
      const float fv[4] = {f4.x, f4.y, f4.z, f4.w};
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        acc[p][0] = fmaf(fv[p], w0.x, acc[p][0]); acc[p][1] = fmaf(fv[p], w0.y, acc[p][1]);
        acc[p][2] = fmaf(fv[p], w0.z, acc[p][2]); acc[p][3] = fmaf(fv[p], w0.w, acc[p][3]);
        acc[p][4] = fmaf(fv[p], w1.x, acc[p][4]); acc[p][5] = fmaf(fv[p], w1.y, acc[p][5]);
        acc[p][6] = fmaf(fv[p], w1.z, acc[p][6]); acc[p][7] = fmaf(fv[p], w1.w, acc[p][7]);
      }
    }
    // ---- ReLU, LayerNorm, NHWC line
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaxf(acc[p][i], 0.f);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += v[i];
      const float mean = warp_sum(s) * (1.0f / D);
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; q += d * d; }
      const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / D) + LN_EPS);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd * gv[i] + ev[i];
      TO* o = out + ((size_t)b * HW + px0 + p0 + p) * 256;
      if (sizeof(TO) == 4) {
        float* of = reinterpret_cast<float*>(o);
        *reinterpret_cast<float4*>(of + lane * 4) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(of + 128 + lane * 4) = make_float4(v[4], v[5], v[6], v[7]);
      } else {
        __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(o);
        __nv_bfloat162 a0 = __floats2bfloat162_rn(v[0], v[1]), a1 = __floats2bfloat162_rn(v[2], v[3]);
        __nv_bfloat162 a2 = __floats2bfloat162_rn(v[4], v[5]), a3 = __floats2bfloat162_rn(v[6], v[7]);
        uint2 lo, hi;
        lo.x = *reinterpret_cast<uint32_t*>(&a0); lo.y = *reinterpret_cast<uint32_t*>(&a1);
        hi.x = *reinterpret_cast<uint32_t*>(&a2); hi.y = *reinterpret_cast<uint32_t*>(&a3);
        *reinterpret_cast<uint2*>(ob + lane * 4) = lo;
        *reinterpret_cast<uint2*>(ob + 128 + lane * 4) = hi;
      }
    }
  }
}

}  // namespace

size_t bev_producer_scratch_bytes(int B, int g, int cb) {
  return ((size_t)B * g * g * 256 + (size_t)cb * 256) * sizeof(float);
}

int launch_bev_producer(const float* tok, const float* map, const float* w, const float* bias,
                        const float* ln_g, const float* ln_b, void* out, int out_bf16, int B, int H, int W,
                        int g, int cb, float* scratch, cudaStream_t st) {
  float* Pt = scratch;
  float* WbT = scratch + (size_t)B * g * g * 256;
  const int n_tok = g * g;
  dim3 grid1((n_tok + 7) / 8, B);
  producer_tokens_kernel<<<grid1, 256, 0, st>>>(tok, w, Pt, WbT, n_tok, 256 + cb, cb);
  const int smem = (cb * 256 + cb * 32) * 4;
  const int n_runs = B * (H * W / 32);
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  const int grid2 = n_runs < 2 * num_sms ? n_runs : 2 * num_sms;
  cudaError_t e;
  if (out_bf16) {
    e = cudaFuncSetAttribute(producer_pixels_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    producer_pixels_kernel<__nv_bfloat16><<<grid2, 256, smem, st>>>(Pt, WbT, map, bias, ln_g, ln_b,
                                                                    reinterpret_cast<__nv_bfloat16*>(out), H, W, g, cb, n_runs);
  } else {
    e = cudaFuncSetAttribute(producer_pixels_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    producer_pixels_kernel<float><<<grid2, 256, smem, st>>>(Pt, WbT, map, bias, ln_g, ln_b,
                                                            reinterpret_cast<float*>(out), H, W, g, cb, n_runs);
  }
  return (int)cudaGetLastError();
}

}  // namespace ddh
