// Shared device helpers of the ddh kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ddh {

constexpr int D = 256;          // d_model == BEV channels == GEMM N-block
constexpr float LN_EPS = 1e-5f;  // nn.LayerNorm default

// ---------------------------------------------------------------------------------
// Row epilogue shared by every GEMM-shaped stage.  One warp owns one output row of a
// 256-wide column block; lane `l` holds the 8 columns  l*4+{0..3}  and  128+l*4+{0..3}
// (so that both the shared-memory reads and the global writes are 16-byte vectors,
// contiguous across the warp).
//
// Order of operations (each optional), chosen so that one descriptor covers every use
// in CustomTransformerDecoderLayer.forward (transfuser_model_v2.py:343-382):
//   acc + bias -> ReLU -> + residual[m] -> LN1 -> + rowvec[m / rows_per_group] -> LN2
//   -> FiLM (x*(1+scale)+shift) -> dot(row, dot_w)+dot_b -> store f32 / bf16
// ---------------------------------------------------------------------------------
struct RowEpi {
  const float* bias = nullptr;
  int relu = 0;
  const float* res = nullptr;
  int ldres = 0;
  const float* ln1_g = nullptr;
  const float* ln1_b = nullptr;
  const float* rowvec = nullptr;
  int rows_per_group = 1;
  const float* ln2_g = nullptr;
  const float* ln2_b = nullptr;
  const float* film = nullptr;   // [2*D]: scale | shift
  const float* dot_w = nullptr;  // [D]
  const float* dot_b = nullptr;  // [1]
  float* dot_out = nullptr;      // [M]
  float* out_f32 = nullptr;
  int ldo32 = 0;
  __nv_bfloat16* out_bf16 = nullptr;
  int ldo16 = 0;
};

__device__ __forceinline__ int lane_col(int lane, int i) {
  return (i < 4) ? (lane * 4 + i) : (128 + lane * 4 + (i - 4));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void layer_norm_row(float (&v)[8], const float* g, const float* b,
                                               int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float d = v[i] - mean;
    q += d * d;
  }
  const float var = warp_sum(q) * (1.0f / D);
  const float rstd = 1.0f / sqrtf(var + LN_EPS);
  const float4 g0 = *reinterpret_cast<const float4*>(g + lane * 4);
  const float4 g1 = *reinterpret_cast<const float4*>(g + 128 + lane * 4);
  const float4 b0 = *reinterpret_cast<const float4*>(b + lane * 4);
  const float4 b1 = *reinterpret_cast<const float4*>(b + 128 + lane * 4);
  const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd * gg[i] + bb[i];
}

__device__ __forceinline__ void load8(const float* p, int lane, float (&o)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p + lane * 4);
  const float4 b = *reinterpret_cast<const float4*>(p + 128 + lane * 4);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
  o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}

// v: the 8 accumulator values of this lane for row m, column block starting at n0.
__device__ __forceinline__ void row_epilogue(const RowEpi& e, float (&v)[8], long long m, int n0,
                                             int lane) {
  float t[8];
  if (e.bias) {
    load8(e.bias + n0, lane, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += t[i];
  }
  if (e.relu) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  if (e.res) {
    load8(e.res + m * e.ldres + n0, lane, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += t[i];
  }
  if (e.ln1_g) layer_norm_row(v, e.ln1_g, e.ln1_b, lane);
  if (e.rowvec) {
    load8(e.rowvec + (m / e.rows_per_group) * D, lane, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += t[i];
  }
  if (e.ln2_g) layer_norm_row(v, e.ln2_g, e.ln2_b, lane);
  if (e.film) {
    float sc[8], sh[8];
    load8(e.film, lane, sc);
    load8(e.film + D, lane, sh);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = v[i] * (1.0f + sc[i]) + sh[i];
  }
  if (e.dot_w) {
    load8(e.dot_w, lane, t);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i] * t[i];
    s = warp_sum(s);
    if (lane == 0) e.dot_out[m] = s + e.dot_b[0];
  }
  if (e.out_f32) {
    float* o = e.out_f32 + m * e.ldo32 + n0;
    *reinterpret_cast<float4*>(o + lane * 4) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(o + 128 + lane * 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  if (e.out_bf16) {
    __nv_bfloat16* o = e.out_bf16 + m * e.ldo16 + n0;
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 p1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]);
    __nv_bfloat162 p3 = __floats2bfloat162_rn(v[6], v[7]);
    uint2 lo, hi;
    lo.x = *reinterpret_cast<uint32_t*>(&p0);
    lo.y = *reinterpret_cast<uint32_t*>(&p1);
    hi.x = *reinterpret_cast<uint32_t*>(&p2);
    hi.y = *reinterpret_cast<uint32_t*>(&p3);
    *reinterpret_cast<uint2*>(o + lane * 4) = lo;
    *reinterpret_cast<uint2*>(o + 128 + lane * 4) = hi;
  }
}

// Mish(x) = x * tanh(softplus(x)), softplus threshold 20 as in torch (nn.Mish / F.mish).
__device__ __forceinline__ float mishf(float x) {
  const float sp = (x > 20.f) ? x : log1pf(expf(x));
  return x * tanhf(sp);
}

// ---------------------------------------------------------------------------------
// Parameters of the GEMM-shaped kernels.  C[m, n0:n0+256] = A[m, :] . W[n0:n0+256, :]^T
// Dense mode: A is [M, K] (fp32 for the SIMT engine, bf16 for the tensor engine).
// Conv  mode: A is gathered on the fly from the NHWC BEV map: row r of scene s is the
//             3x3xC patch around pixel upix[s*rcap + r] (zero padded), K = 9*C, and the
//             output row index is s*rcap + r.  Implements value_proj of
//             GridSampleCrossBEVAttention (modules/blocks.py:68-76,114) evaluated only at
//             the pixels that grid_sample (:117-122) will read.
// ---------------------------------------------------------------------------------
struct GemmParams {
  const void* A = nullptr;
  int lda = 0;
  int M = 0;
  int K = 0;
  const void* W = nullptr;   // SIMT: Wt[K][ldw] fp32; tensor engine: via TMA map
  int ldw = 0;
  int n_blocks = 1;          // 256-wide column blocks of the output (tensor engine)
  RowEpi epi;
  // conv gather
  const void* bev = nullptr;  // NHWC [B][H][W][C]
  const void* bev_lo = nullptr;  // 3xTF32 conv: the low-order plane a - tf32(a) of the fp32 map (bev = the high one)
  const int* upix = nullptr;  // [B][rcap]
  const int* nuniq = nullptr; // [B]
  int rcap = 0;
  int H = 0, W_ = 0, C = 0;
  // fused bilinear/attention combine (tensor-engine conv only)
  const int* ent_slot = nullptr;   // [B][A*ent_per_anchor] slot of the corner pixel or -1
  const float* ent_w = nullptr;    // [B][A*ent_per_anchor] bilinear * attention weight
  int n_anchor = 0;
  int ent_per_anchor = 0;          // P * 4
  long long* dbg = nullptr;        // optional timeline (DDH_TIMELINE builds): CTA 0 clock64 stamps
  int* sched = nullptr;            // tc_conv3_kernel: global scene counter (zeroed before the launch) or null = static deal
  // value rows kept across denoise steps (PlanReuse, kernels.h): bf16 [B * vcap][256] of one layer
  __nv_bfloat16* vout = nullptr;   // tc_conv3_kernel: also store row r of a scene at scene * vcap + r; tc_convv_kernel: output
  int vcap = 0;
  const int2* vrows = nullptr;     // tc_convv_kernel: (pixel index in the batch, value row) per list entry
  const int* n_vrows = nullptr;    // tc_convv_kernel: device count of list entries
};

}  // namespace ddh
