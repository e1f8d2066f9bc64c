// Small-batch ("latency") engine of the ddh planning head: one or two scenes per call.
//
// At batch 1 every GEMM of the decoder chain is 20 x 256 x 256: a 128-row tensor-core tile would
// be 84 % padding and runs on ONE SM.  Here each Linear is split over its OUTPUT COLUMNS across
// many CTAs (8 columns per CTA, one warp per column, lanes split K), activations stay fp32,
// weights are the packed bf16 [N][K] matrices, and LayerNorm / FiLM / the ego add are applied by
// the CONSUMER when it stages its input rows in shared memory ("normalise on load"), so that no
// stage needs a full output row and every stage is a short, wide kernel.
//
// Stored tensors of one decoder layer (transfuser_model_v2.py:343-382), all [M, 256] fp32:
//   x1  = S.Wbo + b + q0                               (S = sum of the conv's per-tile partials)
//   o   = softmax(q.K^T).V,  q = x1.Wq + b             (lat_qattn_kernel, one CTA per scene x head)
//   y2  = o.Wao + b + x1                               (pre-norm1)
//   h   = relu(x2.W1 + b),  x2 = LN2(LN1(y2) + ego)    (prologue LN_EGO_LN)
//   y3  = h.W2 + b                                     (pre-norm3)
//   r1  = relu(x3.Wr0 + b), c1 = relu(x3.Wc0 + b),  x3 = FiLM(LN3(y3))   (prologue LN_FILM)
//   r2  = relu(r1.Wr2 + b), c2 = relu(LN(c1).Wc3 + b)                    (prologue LN)
//   reg = r2.Wr4 + b (reg_finish_kernel),  cls = LN(c2).w6 + b6 (lat_cls_kernel)
#include <stdlib.h>

#include "kernels.h"

namespace ddh {

#ifdef DDH_TIMELINE
__device__ long long g_lat_dbg[64];
#define LI_STAMP(i)                                                                    \
  do {                                                                                 \
    if (threadIdx.x == 0 && blockIdx.x == 0 && p.N == 1024) g_lat_dbg[i] = clock64();  \
  } while (0)
extern "C" __attribute__((visibility("default"))) int ddh_lat_dbg_read(long long* host) {
  return (int)cudaMemcpyFromSymbol(host, g_lat_dbg, sizeof(long long) * 64);
}
#else
#define LI_STAMP(i) do { } while (0)
#endif

enum LatPrologue { LP_NONE = 0, LP_LN = 1, LP_LN_EGO_LN = 2, LP_LN_FILM = 3, LP_SUM_PARTS = 4 };

struct LatLinearParams {
  const float* A = nullptr;        // [M][K] fp32  (LP_SUM_PARTS: [parts][M][K])
  int M = 0, K = 0, N = 0;
  const __nv_bfloat16* W = nullptr;  // [N][K] bf16 (torch Linear layout)
  const float* bias = nullptr;
  int relu = 0;
  const float* res = nullptr;      // [M][N] added after bias / relu
  float* out = nullptr;            // [M][ldo]
  int ldo = 0;
  int prologue = LP_NONE;
  const float* ln1_g = nullptr; const float* ln1_b = nullptr;
  const float* rowvec = nullptr; int rows_per_group = 1;   // LP_LN_EGO_LN
  const float* ln2_g = nullptr; const float* ln2_b = nullptr;
  const float* film = nullptr;                            // LP_LN_FILM: [2*D]
  const int* nuniq = nullptr; int part_stride = 0;         // LP_SUM_PARTS: parts = ceil(nuniq/128)
  int max_parts = 1;
};

__device__ __forceinline__ void lat_ln_row(float (&v)[8], const float* g, const float* b, int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; q += d * d; }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / D) + LN_EPS);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane * 4 + (i & 3) + 128 * (i >> 2);
    v[i] = (v[i] - mean) * rstd * g[c] + b[c];
  }
}

// out[m, n] = epi( pro(A)[m, :] . W[n, :] ),  8 columns per CTA, one warp per column.
// Lane l owns k = 4*l + 128*i + {0..3}: 16-byte shared-memory reads with no bank conflicts and
// 256-byte coalesced weight reads.
struct LatLinearSets { LatLinearParams s[4]; };

// activation loads: read-only path in stand-alone kernels, L1-bypassing (ld.global.cg) inside the
// cooperative layer kernel where other CTAs produced the data earlier in the SAME launch
template <bool CG>
__device__ __forceinline__ float4 ld_act4(const float* p) {
  return CG ? __ldcg(reinterpret_cast<const float4*>(p)) : __ldg(reinterpret_cast<const float4*>(p));
}
template <bool CG>
__device__ __forceinline__ void load8_act(const float* p, int lane, float (&o)[8]) {
  const float4 a = ld_act4<CG>(p + lane * 4);
  const float4 b = ld_act4<CG>(p + 128 + lane * 4);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
  o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}

// global -> shared copy with 8 independent 16-byte loads in flight per thread (a plain
// load-then-store loop serialises one L2 round trip per iteration)
template <bool CG>
__device__ __forceinline__ void stage_copy(float* dst, const float* src, int n4) {
  float4* d = reinterpret_cast<float4*>(dst);
  for (int base = threadIdx.x; base < n4; base += 256 * 8) {
    float4 t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + u * 256;
      if (i < n4) t[u] = ld_act4<CG>(src + (size_t)i * 4);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + u * 256;
      if (i < n4) d[i] = t[u];
    }
  }
}

// One 8-column block of one linear: out[m, n] = epi( pro(A)[m, :] . W[n, :] ), one warp per
// column.  Lane l owns k = 4*l + 128*i + {0..3}: 16-byte shared-memory reads with no bank
// conflicts and 256-byte coalesced weight reads.  Shared memory: rows [M][K].
//
// Measured with clock64 stamps (tools/timeline_lat.py): one item costs ~15 k cycles at batch 1
// and is dominated by cold-start effects (every instruction of a stage executes once per launch)
// rather than by arithmetic; two rewrites that removed the shuffle chains (8 threads per row
// LayerNorm, one thread per output) were SLOWER (more distinct instructions) and were dropped.
template <int KMAX, bool CG>
__device__ __forceinline__ void lat_linear_item(const LatLinearParams& p, int colblock, float* As) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = p.M, K = p.K;
  const int n = colblock * 8 + warp;
  // ---- this warp's weight row: issued first so its latency hides under the input staging
  constexpr int NI = KMAX / 128;
  uint2 wraw[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    wraw[i] = make_uint2(0u, 0u);
    if (i * 128 < K && n < p.N)
      wraw[i] = __ldg(reinterpret_cast<const uint2*>(p.W + (size_t)n * K + i * 128 + lane * 4));
  }
  const float bias = (p.bias && n < p.N) ? __ldg(p.bias + n) : 0.f;
  // ---- stage (and normalise) the input rows
  if (p.prologue == LP_NONE) {
    stage_copy<CG>(As, p.A, M * K / 4);
  } else {
    // K == D for every normalised input; warp w owns rows w, w+8, ...; all of a warp's row
    // loads are issued before any of them is consumed
    constexpr int RMAX = 6;   // M <= 48
    float v[RMAX][8];
    if (p.prologue == LP_SUM_PARTS) {
      int parts[RMAX];
#pragma unroll
      for (int r = 0; r < RMAX; ++r) {
        const int m = warp + 8 * r;
        parts[r] = (m < M) ? (p.nuniq[m / p.rows_per_group] + 127) / 128 : 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[r][i] = 0.f;
      }
      for (int t = 0; t < 8; ++t) {
        bool any = false;
        float u[RMAX][8];
#pragma unroll
        for (int r = 0; r < RMAX; ++r) {
          if (t < parts[r]) {
            any = true;
            load8_act<CG>(p.A + (size_t)t * p.part_stride + (size_t)(warp + 8 * r) * D, lane, u[r]);
          }
        }
        if (!any) break;
#pragma unroll
        for (int r = 0; r < RMAX; ++r)
          if (t < parts[r])
#pragma unroll
            for (int i = 0; i < 8; ++i) v[r][i] += u[r][i];
      }
    } else {
      float e[RMAX][8];
#pragma unroll
      for (int r = 0; r < RMAX; ++r) {
        const int m = warp + 8 * r;
        if (m < M) {
          load8_act<CG>(p.A + (size_t)m * D, lane, v[r]);
          if (p.prologue == LP_LN_EGO_LN)
            load8_act<CG>(p.rowvec + (size_t)(m / p.rows_per_group) * D, lane, e[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < RMAX; ++r) {
        const int m = warp + 8 * r;
        if (m < M) {
          lat_ln_row(v[r], p.ln1_g, p.ln1_b, lane);
          if (p.prologue == LP_LN_EGO_LN) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[r][i] += e[r][i];
            lat_ln_row(v[r], p.ln2_g, p.ln2_b, lane);
          } else if (p.prologue == LP_LN_FILM) {
            float sc[8], sh[8];
            load8(p.film, lane, sc);
            load8(p.film + D, lane, sh);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[r][i] = v[r][i] * (1.0f + sc[i]) + sh[i];
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
      const int m = warp + 8 * r;
      if (m < M) {
        *reinterpret_cast<float4*>(As + m * D + lane * 4) = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
        *reinterpret_cast<float4*>(As + m * D + 128 + lane * 4) = make_float4(v[r][4], v[r][5], v[r][6], v[r][7]);
      }
    }
  }
  __syncthreads();
  if (n >= p.N) return;
  float w[NI][4];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&wraw[i].x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&wraw[i].y);
    w[i][0] = __low2float(a); w[i][1] = __high2float(a);
    w[i][2] = __low2float(b); w[i][3] = __high2float(b);
  }
  // four rows at a time: four independent shuffle reductions in flight
  for (int m0 = 0; m0 < M; m0 += 4) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int m = min(m0 + r, M - 1);
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        if (i * 128 < K) {
          const float4 a = *reinterpret_cast<const float4*>(As + (size_t)m * K + i * 128 + lane * 4);
          s[r] = fmaf(a.x, w[i][0], s[r]);
          s[r] = fmaf(a.y, w[i][1], s[r]);
          s[r] = fmaf(a.z, w[i][2], s[r]);
          s[r] = fmaf(a.w, w[i][3], s[r]);
        }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
      for (int r = 0; r < 4; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], off);
    }
    if (lane < 4 && m0 + lane < M) {
      const int m = m0 + lane;
      float v = (lane == 0) ? s[0] : (lane == 1) ? s[1] : (lane == 2) ? s[2] : s[3];
      v += bias;
      if (p.relu) v = fmaxf(v, 0.f);
      if (p.res) v += CG ? __ldcg(p.res + (size_t)m * p.N + n) : p.res[(size_t)m * p.N + n];
      p.out[(size_t)m * p.ldo + n] = v;
    }
  }
}

template <int KMAX>
__global__ void __launch_bounds__(256) lat_linear_kernel(const LatLinearSets sets) {
  const LatLinearParams& p = sets.s[blockIdx.y];
  if ((int)blockIdx.x * 8 >= p.N) return;
  extern __shared__ __align__(16) float As[];   // [M][K]
  lat_linear_item<KMAX, false>(p, blockIdx.x, As);
}

// q projection + attention core for one (scene, head): q_h = x1.Wq_h + b (32 columns), then
// softmax(q*scale . K_h^T) . V_h over the hoisted K|V rows (nn.MultiheadAttention, :316-321).
// 8 warps: warp w computes q columns 4w..4w+3 (lanes split K), then queries w, w+8, w+16.
template <bool CG>
__device__ __forceinline__ void lat_qattn_item(const float* __restrict__ x1,
                                               const __nv_bfloat16* __restrict__ wq,
                                               const float* __restrict__ bq,
                                               const float* __restrict__ kv, float* __restrict__ o,
                                               int A, int Na, int scene, int h, float* sm) {
  float* As = sm;                       // [A][256]
  float* qs = As + A * D;               // [A][32]
  float* ks = qs + A * 32;              // [Na][33]
  float* vs = ks + Na * 33;             // [Na][32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {
    // K|V head slices first (independent of the producer of x1), 4 keys per thread in flight
    float kt[4], vt[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = tid + u * 256, j = i >> 5, c = i & 31;
      kt[u] = vt[u] = 0.f;
      if (i < Na * 32) {
        kt[u] = __ldg(kv + ((size_t)scene * Na + j) * 2 * D + h * 32 + c);
        vt[u] = __ldg(kv + ((size_t)scene * Na + j) * 2 * D + D + h * 32 + c);
      }
    }
    stage_copy<CG>(As, x1 + (size_t)scene * A * D, A * D / 4);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = tid + u * 256, j = i >> 5, c = i & 31;
      if (i < Na * 32) { ks[j * 33 + c] = kt[u]; vs[j * 32 + c] = vt[u]; }
    }
  }
  __syncthreads();
  const float scale = 0.17677669529663687f;
  {
    // warp w -> q columns 4w..4w+3 of this head; 4 columns x 4 rows of dot products in flight
    float w[4][2][4];
    float bias[4];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int n = h * 32 + warp * 4 + cc;
      bias[cc] = __ldg(bq + n);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(wq + (size_t)n * D + i * 128 + lane * 4));
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
        const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
        w[cc][i][0] = __low2float(a); w[cc][i][1] = __high2float(a);
        w[cc][i][2] = __low2float(b); w[cc][i][3] = __high2float(b);
      }
    }
    for (int m0 = 0; m0 < A; m0 += 4) {
      float s[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int m = min(m0 + r, A - 1);
        const float4 a0 = *reinterpret_cast<const float4*>(As + m * D + lane * 4);
        const float4 a1 = *reinterpret_cast<const float4*>(As + m * D + 128 + lane * 4);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          float t = a0.x * w[cc][0][0];
          t = fmaf(a0.y, w[cc][0][1], t); t = fmaf(a0.z, w[cc][0][2], t); t = fmaf(a0.w, w[cc][0][3], t);
          t = fmaf(a1.x, w[cc][1][0], t); t = fmaf(a1.y, w[cc][1][1], t);
          t = fmaf(a1.z, w[cc][1][2], t); t = fmaf(a1.w, w[cc][1][3], t);
          s[r][cc] = t;
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) s[r][cc] += __shfl_xor_sync(0xffffffffu, s[r][cc], off);
      if (lane < 16) {
        const int r = lane >> 2, cc = lane & 3;
        float v = 0.f;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
#pragma unroll
          for (int c2 = 0; c2 < 4; ++c2) if (rr == r && c2 == cc) v = s[rr][c2];
        if (m0 + r < A) qs[(m0 + r) * 32 + warp * 4 + cc] = (v + bias[cc]) * scale;
      }
    }
  }
  __syncthreads();
  for (int a = warp; a < A; a += 8) {
    float s = -INFINITY;
    if (lane < Na) {
      s = 0.f;
#pragma unroll
      for (int c = 0; c < 32; ++c) s = fmaf(qs[a * 32 + c], ks[lane * 33 + c], s);
    }
    float mx = s;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    const float e = (lane < Na) ? expf(s - mx) : 0.f;
    const float pj = e / warp_sum(e);
    float acc = 0.f;
    for (int j = 0; j < Na; ++j) acc = fmaf(__shfl_sync(0xffffffffu, pj, j), vs[j * 32 + lane], acc);
    o[((size_t)scene * A + a) * D + h * 32 + lane] = acc;
  }
}

__global__ void __launch_bounds__(256) lat_qattn_kernel(const float* __restrict__ x1,
                                                        const __nv_bfloat16* __restrict__ wq,
                                                        const float* __restrict__ bq,
                                                        const float* __restrict__ kv,
                                                        float* __restrict__ o, int A, int Na) {
  extern __shared__ __align__(16) float sm[];
  lat_qattn_item<false>(x1, wq, bq, kv, o, A, Na, blockIdx.x, blockIdx.y, sm);
}

// cls = LN(c2) . w6 + b6, one warp per row (plan_cls_branch tail, :221-224)
template <bool CG>
__device__ __forceinline__ void lat_cls_item(const float* __restrict__ c2, const float* __restrict__ g,
                                             const float* __restrict__ b, const float* __restrict__ w6,
                                             const float* __restrict__ b6, float* __restrict__ scores,
                                             int M, int item) {
  const int m = item * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= M) return;
  float v[8], w[8];
  load8_act<CG>(c2 + (size_t)m * D, lane, v);
  lat_ln_row(v, g, b, lane);
  load8(w6, lane, w);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s = fmaf(v[i], w[i], s);
  s = warp_sum(s);
  if (lane == 0) scores[m] = s + b6[0];
}

__global__ void __launch_bounds__(256) lat_cls_kernel(const float* __restrict__ c2,
                                                      const float* __restrict__ g,
                                                      const float* __restrict__ b,
                                                      const float* __restrict__ w6,
                                                      const float* __restrict__ b6,
                                                      float* __restrict__ scores, int M) {
  lat_cls_item<false>(c2, g, b, w6, b6, scores, M, blockIdx.x);
}

// ---------------------------------------------------------------------------------- launchers
static int lat_linear_smem(const LatLinearArgs& a) { return a.M * a.K * 4; }

static LatLinearParams to_params(const LatLinearArgs& a) {
  LatLinearParams p;
  p.A = a.A; p.M = a.M; p.K = a.K; p.N = a.N; p.W = a.W; p.bias = a.bias; p.relu = a.relu;
  p.res = a.res; p.out = a.out; p.ldo = a.ldo; p.prologue = a.prologue;
  p.ln1_g = a.ln1_g; p.ln1_b = a.ln1_b; p.rowvec = a.rowvec; p.rows_per_group = a.rows_per_group;
  p.ln2_g = a.ln2_g; p.ln2_b = a.ln2_b; p.film = a.film; p.nuniq = a.nuniq;
  p.part_stride = a.part_stride; p.max_parts = a.max_parts;
  return p;
}

// Up to 4 independent linears in ONE launch (grid.y = set): hoisted K|V and ego projections of
// all layers, or the reg/cls branch pairs.
void launch_lat_linear_multi(const LatLinearArgs* args, int n, cudaStream_t st) {
  LatLinearSets sets;
  int smem = 0, kmax = 0, nmax = 0;
  for (int i = 0; i < n; ++i) {
    sets.s[i] = to_params(args[i]);
    const int need = lat_linear_smem(args[i]);
    smem = smem > need ? smem : need;
    kmax = kmax > args[i].K ? kmax : args[i].K;
    nmax = nmax > args[i].N ? nmax : args[i].N;
  }
  static bool once = false;
  if (!once) {
    cudaFuncSetAttribute(lat_linear_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(lat_linear_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(lat_linear_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    once = true;
  }
  dim3 grid((nmax + 7) / 8, n);
  if (kmax <= 256) lat_linear_kernel<256><<<grid, 256, smem, st>>>(sets);
  else if (kmax <= 512) lat_linear_kernel<512><<<grid, 256, smem, st>>>(sets);
  else lat_linear_kernel<1024><<<grid, 256, smem, st>>>(sets);
}

void launch_lat_linear(const LatLinearArgs& a, cudaStream_t st) { launch_lat_linear_multi(&a, 1, st); }

// Regression tail for a handful of rows: CTA = 4 rows, warp p = pose p (outputs 3p..3p+2), lanes
// split K; no shared memory, no block barrier (see reg_finish_kernel for the arithmetic).
template <bool CG>
__device__ __forceinline__ void lat_reg_item(const float* __restrict__ r2, const float* __restrict__ w4,
                                             const float* __restrict__ b4, float* __restrict__ pts,
                                             float* __restrict__ img, float* __restrict__ modes, int M,
                                             int P, int do_ddim, DdimCoef dc, int item) {
  const int lane = threadIdx.x & 31, pose = threadIdx.x >> 5;
  if (pose >= P) return;
  const int m0 = item * 4;
  float w[3][8], s[4][3];
#pragma unroll
  for (int c = 0; c < 3; ++c) load8(w4 + (size_t)(pose * 3 + c) * D, lane, w[c]);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float a[8];
    load8_act<CG>(r2 + (size_t)min(m0 + r, M - 1) * D, lane, a);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t = fmaf(a[i], w[c][i], t);
      s[r][c] = t;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) s[r][c] += __shfl_xor_sync(0xffffffffu, s[r][c], off);
  if (lane < 12) {
    const int r = lane / 3, comp = lane - r * 3, m = m0 + r;
    float mine = 0.f;
#pragma unroll
    for (int rr = 0; rr < 4; ++rr)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) if (rr == r && cc == comp) mine = s[rr][cc];
    if (m < M) {
      mine += b4[pose * 3 + comp];
      float out;
      if (comp < 2) {
        const size_t pi = ((size_t)m * P + pose) * 2 + comp;
        out = __fadd_rn(mine, pts[pi]);
        pts[pi] = out;
        if (do_ddim) {   // norm_odo + DDIM step, same operation order as reg_finish_kernel
          const float x0 = comp ? __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fadd_rn(out, 20.0f)), 46.0f), 1.0f)
                                : __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fadd_rn(out, 1.2f)), 56.9f), 1.0f);
          const float sample = img[pi];
          const float eps = __fdiv_rn(__fsub_rn(sample, __fmul_rn(dc.sqrt_ac_t, x0)), dc.sqrt_1m_ac_t);
          const float x0c = fminf(fmaxf(x0, -1.0f), 1.0f);
          img[pi] = __fadd_rn(__fmul_rn(dc.sqrt_ac_prev, x0c), __fmul_rn(dc.sqrt_1m_ac_prev, eps));
        }
      } else {
        out = __fmul_rn(tanhf(mine), 3.14159265358979323846f);
      }
      modes[((size_t)m * P + pose) * 3 + comp] = out;
    }
  }
}

__global__ void __launch_bounds__(256) lat_reg_finish_kernel(const float* __restrict__ r2,
                                                             const float* __restrict__ w4,
                                                             const float* __restrict__ b4,
                                                             float* __restrict__ pts,
                                                             float* __restrict__ img,
                                                             float* __restrict__ modes, int M, int P,
                                                             int do_ddim, DdimCoef dc) {
  lat_reg_item<false>(r2, w4, b4, pts, img, modes, M, P, do_ddim, dc, blockIdx.x);
}

// ===================================================================================
// One decoder layer after the conv as ONE cooperative kernel: the eight dependent stages
// (output_proj, q-proj + attention, attention out-proj, FFN up, FFN down, reg/cls hidden 1,
// reg/cls hidden 2, regression tail / cls) are separated by grid-wide barriers instead of kernel
// boundaries.  Launched with cudaLaunchCooperativeKernel (all CTAs co-resident).
// ===================================================================================
struct LatLayerParams {
  LatLinearParams bev_out, attn_out, ffn0, ffn2, reg0, cls0, reg2, cls3;
  const __nv_bfloat16* wq; const float* bq; const float* kv; float* o;
  int A, Na, B, heads;
  const float* w4; const float* b4; float* pts; float* img; float* modes;
  int M, P, do_ddim; DdimCoef dc;
  int want_cls; const float *cls_g, *cls_b, *w6, *b6; float* scores;
  unsigned int* bar;
  long long* dbg;   // DDH_TIMELINE builds: clock64 stamps of CTA 0 / CTA 100
};

#ifdef DDH_TIMELINE
#define LT_STAMP(i)                                                                         \
  do {                                                                                      \
    if (P.dbg && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == 100))                \
      P.dbg[(blockIdx.x == 0 ? 0 : 40) + (i)] = clock64();                                  \
  } while (0)
#else
#define LT_STAMP(i) do { } while (0)
#endif

__device__ __forceinline__ void lat_grid_sync(unsigned int* bar, unsigned int& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    const unsigned int target = (epoch + 1) * gridDim.x;
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while (v < target);
    __threadfence();
  }
  ++epoch;
  __syncthreads();
}

template <int KMAX>
__device__ __forceinline__ void lat_linear_stage(const LatLinearParams& p, int cta0, float* As) {
  // CTAs cta0, cta0+1, ... take the 8-column blocks of this linear (grid-stride)
  const int nblk = (p.N + 7) / 8;
  const int me = (int)blockIdx.x - cta0;
  if (me < 0) return;
  for (int blk = me; blk < nblk; blk += (int)gridDim.x - cta0) {
    lat_linear_item<KMAX, true>(p, blk, As);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) lat_layer_kernel(const LatLayerParams P) {
  extern __shared__ __align__(16) float sm[];
  unsigned int epoch = 0;
  const int half = (int)gridDim.x / 2;
  LT_STAMP(0);
  lat_linear_stage<256>(P.bev_out, 0, sm);                                   // x1
  LT_STAMP(1);
  lat_grid_sync(P.bar, epoch);
  LT_STAMP(2);
  for (int it = blockIdx.x; it < P.B * P.heads; it += gridDim.x) {             // o
    lat_qattn_item<true>(P.bev_out.out, P.wq, P.bq, P.kv, P.o, P.A, P.Na, it / P.heads, it % P.heads, sm);
    __syncthreads();
  }
  LT_STAMP(3);
  lat_grid_sync(P.bar, epoch);
  LT_STAMP(4);
  lat_linear_stage<256>(P.attn_out, 0, sm);                                  // y2
  LT_STAMP(5);
  lat_grid_sync(P.bar, epoch);
  LT_STAMP(6);
  lat_linear_stage<256>(P.ffn0, 0, sm);                                      // h
  LT_STAMP(7);
  lat_grid_sync(P.bar, epoch);
  LT_STAMP(8);
  lat_linear_stage<1024>(P.ffn2, 0, sm);                                     // y3
  LT_STAMP(9);
  lat_grid_sync(P.bar, epoch);
  LT_STAMP(10);
  lat_linear_stage<256>(P.reg0, 0, sm);                                      // r1 | c1
  if (P.want_cls && (int)blockIdx.x >= half) lat_linear_stage<256>(P.cls0, half, sm);
  LT_STAMP(11);
  lat_grid_sync(P.bar, epoch);
  LT_STAMP(12);
  lat_linear_stage<256>(P.reg2, 0, sm);                                      // r2 | c2
  if (P.want_cls && (int)blockIdx.x >= half) lat_linear_stage<256>(P.cls3, half, sm);
  LT_STAMP(13);
  lat_grid_sync(P.bar, epoch);
  LT_STAMP(14);
  if ((int)blockIdx.x < half) {                                               // tail | cls
    for (int it = blockIdx.x; it * 4 < P.M; it += half)
      lat_reg_item<true>(P.reg2.out, P.w4, P.b4, P.pts, P.img, P.modes, P.M, P.P, P.do_ddim, P.dc, it);
  } else if (P.want_cls) {
    for (int it = (int)blockIdx.x - half; it * 8 < P.M; it += (int)gridDim.x - half)
      lat_cls_item<true>(P.cls3.out, P.cls_g, P.cls_b, P.w6, P.b6, P.scores, P.M, it);
  }
  LT_STAMP(15);
}

int launch_lat_layer(const LatLayerArgs& a, cudaStream_t st) {
  LatLayerParams P;
  P.bev_out = to_params(a.bev_out); P.attn_out = to_params(a.attn_out); P.ffn0 = to_params(a.ffn0);
  P.ffn2 = to_params(a.ffn2); P.reg0 = to_params(a.reg0); P.cls0 = to_params(a.cls0);
  P.reg2 = to_params(a.reg2); P.cls3 = to_params(a.cls3);
  P.wq = a.wq; P.bq = a.bq; P.kv = a.kv; P.o = a.o; P.A = a.A; P.Na = a.Na; P.B = a.B; P.heads = a.heads;
  P.w4 = a.w4; P.b4 = a.b4; P.pts = a.pts; P.img = a.img; P.modes = a.modes;
  P.M = a.M; P.P = a.P; P.do_ddim = a.do_ddim; P.dc = a.dc;
  P.want_cls = a.want_cls; P.cls_g = a.cls_g; P.cls_b = a.cls_b; P.w6 = a.w6; P.b6 = a.b6;
  P.scores = a.scores; P.bar = a.bar; P.dbg = a.dbg;
  int smem = (a.A * D + a.A * 32 + a.Na * 33 + a.Na * 32) * 4;
  const LatLinearArgs* all[8] = {&a.bev_out, &a.attn_out, &a.ffn0, &a.ffn2, &a.reg0, &a.cls0, &a.reg2, &a.cls3};
  for (int i = 0; i < 8; ++i) {
    const int need = lat_linear_smem(*all[i]);
    smem = smem > need ? smem : need;
  }
  static int cur = 0, grid = 0;
  if (smem > cur) {
    cudaFuncSetAttribute(lat_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cur = smem;
  }
  if (!grid) {
    int dev = 0, sms = 0, coop = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop || sms < 16) return 1;
    grid = sms < 128 ? (sms & ~1) : 128;
    if (const char* e = getenv("DDH_LAT_GRID")) {
      const int g = atoi(e);
      if (g >= 16 && g <= sms) grid = g & ~1;
    }
  }
  void* args[] = {&P};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)lat_layer_kernel, dim3(grid), dim3(256),
                                              args, (size_t)smem, st);
  return e == cudaSuccess ? 0 : 2;
}

void launch_lat_reg_finish(const float* r2, const float* w4, const float* b4, float* pts, float* img,
                           float* modes, int M, int P, int do_ddim, DdimCoef dc, cudaStream_t st) {
  lat_reg_finish_kernel<<<(M + 3) / 4, 256, 0, st>>>(r2, w4, b4, pts, img, modes, M, P, do_ddim, dc);
}

void launch_lat_qattn(const float* x1, const __nv_bfloat16* wq, const float* bq, const float* kv,
                      float* o, int B, int A, int Na, int heads, cudaStream_t st) {
  const int smem = (A * D + A * 32 + Na * 33 + Na * 32) * 4;
  static int cur = 0;
  if (smem > cur && smem > 48 * 1024) {
    cudaFuncSetAttribute(lat_qattn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cur = smem;
  }
  dim3 grid(B, heads);
  lat_qattn_kernel<<<grid, 256, smem, st>>>(x1, wq, bq, kv, o, A, Na);
}

void launch_lat_cls(const float* c2, const float* g, const float* b, const float* w6,
                    const float* b6, float* scores, int M, cudaStream_t st) {
  lat_cls_kernel<<<(M + 7) / 8, 256, 0, st>>>(c2, g, b, w6, b6, scores, M);
}

}  // namespace ddh
