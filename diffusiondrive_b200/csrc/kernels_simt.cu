// CUDA-core (fp32) kernels of the ddh planning head: the fp32-precision GEMM/conv engine
// and every non-GEMM stage (layout conversion, sine embedding, BEV sampling plan, bilinear
// combine, agent attention core, regression tail + DDIM, mode selection).
//
// Reference lines are cited per kernel (paths relative to
// navsim/agents/diffusiondrive/ of seulbinHwang/DiffusionDrive).
#include <math.h>

#include "kernels.h"
#include "geom.cuh"
#include "tc_ptx.cuh"

namespace ddh {

// ===================================================================================
// fp32 GEMM engine:  C[m, n0:n0+256] = A[m,:] . Wt[:, n0:n0+256]  (+ fused row epilogue)
// Tile 64 rows x 256 cols x 16 k, 256 threads, 8x8 outputs per thread.
// CONV = true gathers A rows from the NHWC fp32 BEV map (value_proj conv evaluated at the
// sampled pixels only, modules/blocks.py:68-76,114).
// ===================================================================================
constexpr int SG_BM = 64, SG_BK = 16, SG_THREADS = 256;
constexpr int SG_AS_LD = SG_BM + 4;
constexpr int SG_SMEM_BYTES = (SG_BK * SG_AS_LD + SG_BK * D + SG_BM * D) * 4 + SG_BM * 4;

template <bool CONV>
__global__ void __launch_bounds__(SG_THREADS, 2) simt_gemm_kernel(const GemmParams p) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                         // [BK][AS_LD]  (k-major: As[k][row])
  float* Ws = As + SG_BK * SG_AS_LD;        // [BK][256]
  float* Cs = Ws + SG_BK * D;               // [BM][256]
  int* s_pix = reinterpret_cast<int*>(Cs + SG_BM * D);  // [BM] (conv only)

  const int tid = threadIdx.x;
  const int tx = tid & 31, ty = tid >> 5;
  const int n0 = CONV ? 0 : blockIdx.y * D;
  int row0, rows_valid;
  long long out_row0;
  int scene = 0;
  if (CONV) {
    scene = blockIdx.y;
    const int nu = p.nuniq[scene];
    row0 = blockIdx.x * SG_BM;
    if (row0 >= nu) return;
    rows_valid = min(SG_BM, nu - row0);
    out_row0 = (long long)scene * p.rcap + row0;
    if (tid < SG_BM) s_pix[tid] = (tid < rows_valid) ? p.upix[(long long)scene * p.rcap + row0 + tid] : -1;
    __syncthreads();
  } else {
    row0 = blockIdx.x * SG_BM;
    rows_valid = min(SG_BM, p.M - row0);
    out_row0 = row0;
  }

  // A-tile load mapping: thread -> (row = tid/4, 4 consecutive k at (tid%4)*4)
  const int a_row = tid >> 2, a_kq = (tid & 3) * 4;
  const float* Af = reinterpret_cast<const float*>(p.A);
  const float* bev = reinterpret_cast<const float*>(p.bev);
  int a_y = 0, a_x = 0;
  bool a_valid = a_row < rows_valid;
  if (CONV && a_valid) {
    const int yx = s_pix[a_row];   // (y << 16) | x
    a_y = yx >> 16;
    a_x = yx & 0xffff;
  }
  const float* Wt = reinterpret_cast<const float*>(p.W) + n0;

  auto load_a = [&](int k0) -> float4 {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a_valid) {
      if (CONV) {
        const int tap = k0 / p.C, c0 = k0 - tap * p.C;
        const int yy = a_y + tap / 3 - 1, xx = a_x + tap % 3 - 1;
        if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W_) {
          const float* src = bev + (((long long)scene * p.H + yy) * p.W_ + xx) * p.C + c0 + a_kq;
          v = *reinterpret_cast<const float4*>(src);
        }
      } else {
        v = *reinterpret_cast<const float4*>(Af + (long long)(row0 + a_row) * p.lda + k0 + a_kq);
      }
    }
    return v;
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nk = p.K / SG_BK;
  float4 a_reg = load_a(0);
  float4 w_reg[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + i * SG_THREADS;  // float4 index in [BK][64]
    const int k = idx >> 6, c4 = idx & 63;
    w_reg[i] = *reinterpret_cast<const float4*>(Wt + (long long)k * p.ldw + c4 * 4);
  }

  for (int kc = 0; kc < nk; ++kc) {
    // registers -> smem
    As[(a_kq + 0) * SG_AS_LD + a_row] = a_reg.x;
    As[(a_kq + 1) * SG_AS_LD + a_row] = a_reg.y;
    As[(a_kq + 2) * SG_AS_LD + a_row] = a_reg.z;
    As[(a_kq + 3) * SG_AS_LD + a_row] = a_reg.w;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * SG_THREADS;
      *reinterpret_cast<float4*>(Ws + idx * 4) = w_reg[i];
    }
    __syncthreads();
    if (kc + 1 < nk) {  // prefetch next tile into registers
      const int k0 = (kc + 1) * SG_BK;
      a_reg = load_a(k0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = tid + i * SG_THREADS;
        const int k = idx >> 6, c4 = idx & 63;
        w_reg[i] = *reinterpret_cast<const float4*>(Wt + (long long)(k0 + k) * p.ldw + c4 * 4);
      }
    }
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(As + k * SG_AS_LD + ty * 8);
      const float4 a1 = *reinterpret_cast<const float4*>(As + k * SG_AS_LD + ty * 8 + 4);
      const float4 w0 = *reinterpret_cast<const float4*>(Ws + k * D + tx * 4);
      const float4 w1 = *reinterpret_cast<const float4*>(Ws + k * D + 128 + tx * 4);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }

  // stage the tile, then one warp per row runs the shared epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float* c = Cs + (ty * 8 + i) * D;
    *reinterpret_cast<float4*>(c + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    *reinterpret_cast<float4*>(c + 128 + tx * 4) =
        make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
  }
  __syncthreads();
  for (int r = ty; r < rows_valid; r += SG_THREADS / 32) {
    float v[8];
    load8(Cs + r * D, tx, v);
    row_epilogue(p.epi, v, out_row0 + r, n0, tx);
  }
}

void launch_simt_gemm(const GemmParams& p, int n_total, cudaStream_t st) {
  static bool once = false;
  if (!once) {
    cudaFuncSetAttribute(simt_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         SG_SMEM_BYTES);
    cudaFuncSetAttribute(simt_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         SG_SMEM_BYTES);
    once = true;
  }
  dim3 grid((p.M + SG_BM - 1) / SG_BM, n_total / D);
  simt_gemm_kernel<false><<<grid, SG_THREADS, SG_SMEM_BYTES, st>>>(p);
}

void launch_simt_conv(const GemmParams& p, int B, cudaStream_t st) {
  static bool once = false;
  if (!once) {
    cudaFuncSetAttribute(simt_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         SG_SMEM_BYTES);
    cudaFuncSetAttribute(simt_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         SG_SMEM_BYTES);
    once = true;
  }
  dim3 grid((p.rcap + SG_BM - 1) / SG_BM, B);
  simt_gemm_kernel<true><<<grid, SG_THREADS, SG_SMEM_BYTES, st>>>(p);
}

// ===================================================================================
// BEV layout conversion  [B][C][HW] (NCHW)  ->  [B][HW][C] (NHWC), optional dtype change.
// The reference hands the head NCHW fp32 (transfuser_model_v2.py:138-140); the gather of
// the on-demand value_proj wants all channels of one pixel contiguous.
// One CTA moves a 64-pixel x 256-channel tile through shared memory: coalesced 256-byte
// reads along HW, coalesced 128-byte writes along C.  HBM-bound: 4 B read + 2|4 B written
// per element.
// ===================================================================================
template <typename TI>
__device__ __forceinline__ float4 ld4(const TI* p);
template <>
__device__ __forceinline__ float4 ld4<float>(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
template <>
__device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}

// PXT pixels per tile: 64 for fp32 output; 32 for bf16 output, whose 16.5 KB of shared memory
// lets a layout CTA share an SM with the resident tensor-core CTAs (scene-chunk concurrency).
template <typename TI, typename TO, int PXT>
__global__ void __launch_bounds__(256) bev_to_nhwc_kernel(const TI* __restrict__ src,
                                                          TO* __restrict__ dst, int C, int HW) {
  // tile[px][c], padded so that the transposing stores are conflict free
  constexpr int LDW = (sizeof(TO) == 2) ? 129 : 257;  // 32-bit words per pixel row
  extern __shared__ __align__(16) uint32_t tile_u32[];
  TO* tile = reinterpret_cast<TO*>(tile_u32);
  constexpr int LDE = LDW * 4 / sizeof(TO);  // elements per pixel row
  constexpr int G = PXT / 4;                 // groups of 4 pixels
  constexpr int CL = 256 / G;                // channels covered per iteration
  const int b = blockIdx.y;
  const int px0 = blockIdx.x * PXT;
  const int tid = threadIdx.x;
  const int px4 = tid % G;
  const int cl = tid / G;
  const TI* s = src + (size_t)b * C * HW + px0 + px4 * 4;
  float4 v[256 / CL];
#pragma unroll
  for (int i = 0; i < 256 / CL; ++i) v[i] = ld4<TI>(s + (size_t)(i * CL + cl) * HW);
#pragma unroll
  for (int i = 0; i < 256 / CL; ++i) {
    const int c = i * CL + cl;
    tile[(px4 * 4 + 0) * LDE + c] = (TO)v[i].x;
    tile[(px4 * 4 + 1) * LDE + c] = (TO)v[i].y;
    tile[(px4 * 4 + 2) * LDE + c] = (TO)v[i].z;
    tile[(px4 * 4 + 3) * LDE + c] = (TO)v[i].w;
  }
  __syncthreads();
  // write: one warp per pixel row, 32-bit words, consecutive lanes consecutive words
  const int lane = tid & 31, warp = tid >> 5;
  constexpr int WORDS = 256 * sizeof(TO) / 4;  // 128 (bf16) or 256 (f32)
  uint32_t* d = reinterpret_cast<uint32_t*>(dst + ((size_t)b * HW + px0) * C);
  for (int px = warp; px < PXT; px += 8) {
#pragma unroll
    for (int w = lane; w < WORDS; w += 32) d[(size_t)px * WORDS + w] = tile_u32[px * LDW + w];
  }
}

template <typename TI, typename TO>
static void bev_launch(const void* src, void* dst, int B, int C, int HW, cudaStream_t st) {
  constexpr int LDW = (sizeof(TO) == 2) ? 129 : 257;
  constexpr int PXT = (sizeof(TO) == 2) ? 32 : 64;
  const int smem = PXT * LDW * 4;
  static bool once = false;
  if (!once) {
    cudaFuncSetAttribute(bev_to_nhwc_kernel<TI, TO, PXT>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    once = true;
  }
  dim3 grid(HW / PXT, B);
  bev_to_nhwc_kernel<TI, TO, PXT><<<grid, 256, smem, st>>>(reinterpret_cast<const TI*>(src),
                                                          reinterpret_cast<TO*>(dst), C, HW);
}

// On-demand variant: converts only the BEV segments (SEG pixels of one row) the coming conv call
// reads and that are not converted yet (todo bit mask written by plan_kernel, which also maintains
// the done mask).  One CTA per scene at large batch (segments dealt over gridDim.y CTAs at small
// batch), thread = channel: per step two segments are read (SEG*4 contiguous bytes per channel
// plane), transposed through shared memory and written as SEG x 512-byte pixel lines.
// Trajectories sample ~250 of 4096 pixels; with 8-pixel segments ~20 % of the map is converted.
template <typename TI, typename TO, int SEG, int NSEG>
__global__ void __launch_bounds__(256) bev_segs_to_nhwc_kernel(
    const TI* __restrict__ src, TO* __restrict__ dst, const unsigned int* __restrict__ todo,
    int nw32, int C, int H, int W, TO* __restrict__ dst_lo) {
  extern __shared__ __align__(16) unsigned char seg_raw[];   // NSEG segments per step
  TO* tile = reinterpret_cast<TO*>(seg_raw);    // [NSEG * SEG px][256 channels]
  __shared__ unsigned short list[2048];
  __shared__ int cnt[65];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int HW = H * W, segs_per_row = W / SEG;
  unsigned int word = 0;
  if (tid < nw32) word = todo[(size_t)b * nw32 + tid];
  if (tid < 64) cnt[tid] = __popc(word);
  __syncthreads();
  if (tid < nw32) {
    int off = 0;
    for (int i = 0; i < tid; ++i) off += cnt[i];
    while (word) {
      const int bit = __ffs((int)word) - 1;
      word &= word - 1;
      list[off++] = (unsigned short)(tid * 32 + bit);
    }
    if (tid == nw32 - 1) cnt[64] = off;
  }
  __syncthreads();
  const int n = cnt[64];
  if (n == 0) return;
  const TI* sb = src + (size_t)b * C * HW + (size_t)tid * HW;
  TO* db = dst + (size_t)b * HW * C;
  constexpr int V4_PER_ROW = 256 * sizeof(TO) / 16;   // uint4 per pixel line
  for (int i = NSEG * blockIdx.y; i < n; i += NSEG * gridDim.y) {
    const int ns = min(NSEG, n - i);
    int px0[NSEG];
    float4 v[NSEG][SEG / 4];
#pragma unroll
    for (int sgi = 0; sgi < NSEG; ++sgi) {
      const int sidx = list[min(i + sgi, n - 1)];
      const int y = sidx / segs_per_row;
      px0[sgi] = y * W + (sidx - y * segs_per_row) * SEG;
      if (sgi < ns) {
#pragma unroll
        for (int q = 0; q < SEG / 4; ++q) v[sgi][q] = ld4<TI>(sb + px0[sgi] + 4 * q);
      }
    }
    __syncthreads();   // previous step fully written out
#pragma unroll
    for (int sgi = 0; sgi < NSEG; ++sgi) {
      if (sgi < ns) {
#pragma unroll
        for (int q = 0; q < SEG / 4; ++q) {
          TO* t = tile + (size_t)(sgi * SEG + 4 * q) * 256 + tid;
          t[0] = (TO)v[sgi][q].x; t[256] = (TO)v[sgi][q].y; t[512] = (TO)v[sgi][q].z; t[768] = (TO)v[sgi][q].w;
        }
      }
    }
    __syncthreads();
    for (int u = tid; u < ns * SEG * V4_PER_ROW; u += 256) {
      const int row = u / V4_PER_ROW, part = u - row * V4_PER_ROW;
      const int sgi = row / SEG, px = row - sgi * SEG;
      const uint4 v = reinterpret_cast<const uint4*>(tile + (size_t)row * 256)[part];
      if (sizeof(TO) == 4 && dst_lo) {
        // fp32 engine with the 3xTF32 conv: high plane = the 10 mantissa bits the tensor core reads,
        // low plane = the exact remainder
        const uint4 hi = make_uint4(v.x & 0xFFFFE000u, v.y & 0xFFFFE000u, v.z & 0xFFFFE000u, v.w & 0xFFFFE000u);
        const uint4 lo = make_uint4(__float_as_uint(__uint_as_float(v.x) - __uint_as_float(hi.x)),
                                    __float_as_uint(__uint_as_float(v.y) - __uint_as_float(hi.y)),
                                    __float_as_uint(__uint_as_float(v.z) - __uint_as_float(hi.z)),
                                    __float_as_uint(__uint_as_float(v.w) - __uint_as_float(hi.w)));
        reinterpret_cast<uint4*>(db + (size_t)(px0[sgi] + px) * C)[part] = hi;
        reinterpret_cast<uint4*>(dst_lo + (size_t)b * HW * C + (size_t)(px0[sgi] + px) * C)[part] = lo;
      } else {
        reinterpret_cast<uint4*>(db + (size_t)(px0[sgi] + px) * C)[part] = v;
      }
    }
  }
}

// 32-pixel segments with lanes along the pixels: 8 consecutive lanes read 128 contiguous bytes of one
// channel plane.  Used when the source map is pinned HOST memory read in place across PCIe
// (ddh_forward_host): requests of 128 bytes instead of 32 keep the link busy.
template <typename TI, typename TO, int PXT>
__global__ void __launch_bounds__(256) bev_segs_px_to_nhwc_kernel(
    const TI* __restrict__ src, TO* __restrict__ dst, const unsigned int* __restrict__ todo,
    int nw32, int C, int H, int W) {
  constexpr int LDW = (sizeof(TO) == 2) ? 129 : 257;
  constexpr int G = PXT / 4;          // lanes along the pixels (float4 each)
  constexpr int CL = 256 / G;         // channels covered per iteration
  extern __shared__ __align__(16) unsigned char seg_raw[];
  uint32_t* tile_u32 = reinterpret_cast<uint32_t*>(seg_raw);
  TO* tile = reinterpret_cast<TO*>(seg_raw);
  constexpr int LDE = LDW * 4 / sizeof(TO);
  __shared__ unsigned short list[2048];
  __shared__ int cnt[65];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int HW = H * W, segs_per_row = W / PXT;
  unsigned int word = 0;
  if (tid < nw32) word = todo[(size_t)b * nw32 + tid];
  if (tid < 64) cnt[tid] = __popc(word);
  __syncthreads();
  if (tid < nw32) {
    int off = 0;
    for (int i = 0; i < tid; ++i) off += cnt[i];
    while (word) {
      const int bit = __ffs((int)word) - 1;
      word &= word - 1;
      list[off++] = (unsigned short)(tid * 32 + bit);
    }
    if (tid == nw32 - 1) cnt[64] = off;
  }
  __syncthreads();
  const int n = cnt[64];
  const int px4 = tid % G, cl = tid / G, lane = tid & 31, warp = tid >> 5;
  constexpr int WORDS = 256 * sizeof(TO) / 4;
  for (int i = blockIdx.y; i < n; i += gridDim.y) {
    const int sidx = list[i];
    const int y = sidx / segs_per_row;
    const int px0 = y * W + (sidx - y * segs_per_row) * PXT;
    const TI* sp = src + (size_t)b * C * HW + px0 + px4 * 4;
    float4 v[256 / CL];
#pragma unroll
    for (int k = 0; k < 256 / CL; ++k) v[k] = ld4<TI>(sp + (size_t)(k * CL + cl) * HW);
    __syncthreads();   // previous tile fully written out
#pragma unroll
    for (int k = 0; k < 256 / CL; ++k) {
      const int c = k * CL + cl;
      tile[(px4 * 4 + 0) * LDE + c] = (TO)v[k].x;
      tile[(px4 * 4 + 1) * LDE + c] = (TO)v[k].y;
      tile[(px4 * 4 + 2) * LDE + c] = (TO)v[k].z;
      tile[(px4 * 4 + 3) * LDE + c] = (TO)v[k].w;
    }
    __syncthreads();
    uint32_t* d = reinterpret_cast<uint32_t*>(dst + ((size_t)b * HW + px0) * C);
    for (int px = warp; px < PXT; px += 8) {
#pragma unroll
      for (int w = lane; w < WORDS; w += 32) d[(size_t)px * WORDS + w] = tile_u32[px * LDW + w];
    }
  }
}

template <typename TI, typename TO>
static void bev_segs_launch(const void* src, void* dst, const unsigned int* todo, int nw32, int seg,
                            int B, int C, int H, int W, cudaStream_t st, void* dst_lo = nullptr) {
  int ysplit = 592 / (B > 0 ? B : 1);     // small batches: spread one scene's segments over CTAs
  ysplit = ysplit < 1 ? 1 : (ysplit > 32 ? 32 : ysplit);
  dim3 grid(B, ysplit);
  // 8-pixel segments: four per step (consecutive list entries are usually neighbours in x, so a
  // thread's reads of one channel plane coalesce into 64/128-byte runs); 16-pixel: two per step
  if (seg == 64) {
    constexpr int smem64 = 64 * ((sizeof(TO) == 2) ? 129 : 257) * 4;
    static bool once = false;
    if (!once) {
      cudaFuncSetAttribute(bev_segs_px_to_nhwc_kernel<TI, TO, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem64);
      once = true;
    }
    bev_segs_px_to_nhwc_kernel<TI, TO, 64><<<grid, 256, smem64, st>>>(
        reinterpret_cast<const TI*>(src), reinterpret_cast<TO*>(dst), todo, nw32, C, H, W);
  } else if (seg == 32)
    bev_segs_px_to_nhwc_kernel<TI, TO, 32><<<grid, 256, 32 * ((sizeof(TO) == 2) ? 129 : 257) * 4, st>>>(
        reinterpret_cast<const TI*>(src), reinterpret_cast<TO*>(dst), todo, nw32, C, H, W);
  else if (seg == -16)   // 16-pixel segments, lanes along the pixels (pinned host source)
    bev_segs_px_to_nhwc_kernel<TI, TO, 16><<<grid, 256, 16 * ((sizeof(TO) == 2) ? 129 : 257) * 4, st>>>(
        reinterpret_cast<const TI*>(src), reinterpret_cast<TO*>(dst), todo, nw32, C, H, W);
  else if (seg == 8)
    bev_segs_to_nhwc_kernel<TI, TO, 8, 4><<<grid, 256, 4 * 8 * 256 * sizeof(TO), st>>>(
        reinterpret_cast<const TI*>(src), reinterpret_cast<TO*>(dst), todo, nw32, C, H, W, reinterpret_cast<TO*>(dst_lo));
  else
    bev_segs_to_nhwc_kernel<TI, TO, 16, 2><<<grid, 256, 2 * 16 * 256 * sizeof(TO), st>>>(
        reinterpret_cast<const TI*>(src), reinterpret_cast<TO*>(dst), todo, nw32, C, H, W, reinterpret_cast<TO*>(dst_lo));
}

void launch_bev_segs_to_nhwc(const void* src, int src_dtype, void* dst, int dst_dtype,
                             const unsigned int* todo, int nw32, int seg, int B, int C, int H, int W,
                             cudaStream_t st, void* dst_lo) {
  // dst_lo (fp32 destination, 8- or 16-pixel device segments only): write the map as a high / low
  // plane pair for the 3xTF32 conv
  if (src_dtype == 0 && dst_dtype == 0) bev_segs_launch<float, float>(src, dst, todo, nw32, seg, B, C, H, W, st, dst_lo);
  else if (src_dtype == 0 && dst_dtype == 1)
    bev_segs_launch<float, __nv_bfloat16>(src, dst, todo, nw32, seg, B, C, H, W, st);
  else if (src_dtype == 1 && dst_dtype == 1)
    bev_segs_launch<__nv_bfloat16, __nv_bfloat16>(src, dst, todo, nw32, seg, B, C, H, W, st);
  else bev_segs_launch<__nv_bfloat16, float>(src, dst, todo, nw32, seg, B, C, H, W, st, dst_lo);
}

void launch_bev_to_nhwc(const void* src, int src_dtype, void* dst, int dst_dtype, int B, int C,
                        int HW, cudaStream_t st) {
  if (src_dtype == 0 && dst_dtype == 0) bev_launch<float, float>(src, dst, B, C, HW, st);
  else if (src_dtype == 0 && dst_dtype == 1) bev_launch<float, __nv_bfloat16>(src, dst, B, C, HW, st);
  else if (src_dtype == 1 && dst_dtype == 1)
    bev_launch<__nv_bfloat16, __nv_bfloat16>(src, dst, B, C, HW, st);
  else bev_launch<__nv_bfloat16, float>(src, dst, B, C, HW, st);
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ d,
                                     size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(s) + i);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    reinterpret_cast<uint2*>(d)[i] = u;
  }
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ s, float* __restrict__ d,
                                     size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) reinterpret_cast<float4*>(d)[i] = ld4<__nv_bfloat16>(s + i * 4);
}
// 3xTF32 operand planes of an fp32 array: hi = the 10 mantissa bits the tensor core reads, lo = x - hi
// (exact).  src may alias hi.
__global__ void split_tf32_kernel(const float* s, float* hi, float* lo, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(s)[i];
    float4 h;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
    reinterpret_cast<float4*>(hi)[i] = h;
    reinterpret_cast<float4*>(lo)[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
  }
}
void launch_split_tf32(const float* src, float* hi, float* lo, size_t n, cudaStream_t st) {
  const size_t n4 = n / 4;
  const int blocks = (int)min((size_t)148 * 16, (n4 + 255) / 256);
  split_tf32_kernel<<<blocks > 0 ? blocks : 1, 256, 0, st>>>(src, hi, lo, n4);
}
void launch_cast_f32_bf16(const float* src, __nv_bfloat16* dst, size_t n, cudaStream_t st) {
  const size_t n4 = n / 4;
  const int blocks = (int)min((size_t)148 * 16, (n4 + 255) / 256);
  cast_f32_bf16_kernel<<<blocks > 0 ? blocks : 1, 256, 0, st>>>(src, dst, n4);
}
void launch_cast_bf16_f32(const __nv_bfloat16* src, float* dst, size_t n, cudaStream_t st) {
  const size_t n4 = n / 4;
  const int blocks = (int)min((size_t)148 * 16, (n4 + 255) / 256);
  cast_bf16_f32_kernel<<<blocks > 0 ? blocks : 1, 256, 0, st>>>(src, dst, n4);
}

// ===================================================================================
// Odometry (de)normalisation, transfuser_model_v2.py:480-500.  Intrinsics keep the
// reference's operation order (no FMA contraction) so fp32 results track torch's.
// ===================================================================================
// img = sqrt(ac[t]) * norm_odo(plan_anchor) + sqrt(1-ac[t]) * noise     (:591-597)
__global__ void init_img_kernel(const float* __restrict__ anchors, const float* __restrict__ noise,
                                float* __restrict__ img, int B, int AP, float sa, float sb) {
  const size_t n = (size_t)B * AP * 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % ((size_t)AP * 2));
    const float a = anchors[j];
    const float nv = (j & 1) ? norm_y(a) : norm_x(a);
    img[i] = __fadd_rn(__fmul_rn(sa, nv), __fmul_rn(sb, noise[i]));
  }
}
void launch_init_img(const float* anchors, const float* noise, float* img, int B, int AP,
                     float sqrt_ac, float sqrt_1m_ac, cudaStream_t st) {
  const size_t n = (size_t)B * AP * 2;
  const int blocks = (int)min((size_t)148 * 8, (n + 255) / 256);
  init_img_kernel<<<blocks, 256, 0, st>>>(anchors, noise, img, B, AP, sqrt_ac, sqrt_1m_ac);
}

// ===================================================================================
// clamp + denorm_odo (:601-602) and gen_sineembed_for_position(hidden_dim=64)
// (modules/blocks.py:22-40, call :605-607): per pose 64 features = [embed(y) | embed(x)],
// embed(v)[i] = sin|cos(v * 2pi / dim_t[i]) (even i: sin, odd i: cos), flattened over the
// P poses -> 64*P features per anchor row.  Accurate sinf/cosf: arguments reach ~360 rad.
// 64 threads per row, 4 rows per CTA.
// ===================================================================================
__global__ void __launch_bounds__(256) embed_kernel(const float* __restrict__ img,
                                                    float* __restrict__ pts,
                                                    float* __restrict__ emb32,
                                                    __nv_bfloat16* __restrict__ emb16, int M, int P,
                                                    const float* __restrict__ dim_t) {
  const int m = blockIdx.x * 4 + (threadIdx.x >> 6);
  if (m >= M) return;
  const int j = threadIdx.x & 63;
  const int half = j >> 5, i = j & 31;
  const float dt = dim_t[i];
  const float two_pi = 6.283185307179586f;
  for (int p = 0; p < P; ++p) {
    const float ix = img[((size_t)m * P + p) * 2 + 0];
    const float iy = img[((size_t)m * P + p) * 2 + 1];
    const float x = denorm_x(fminf(fmaxf(ix, -1.0f), 1.0f));
    const float y = denorm_y(fminf(fmaxf(iy, -1.0f), 1.0f));
    if (j == 0) {
      pts[((size_t)m * P + p) * 2 + 0] = x;
      pts[((size_t)m * P + p) * 2 + 1] = y;
    }
    const float v = half ? x : y;  // output order is (pos_y, pos_x), blocks.py:39
    const float arg = __fdiv_rn(__fmul_rn(v, two_pi), dt);
    const float e = (i & 1) ? cosf(arg) : sinf(arg);
    const size_t o = (size_t)m * (64 * P) + p * 64 + j;
    if (emb32) emb32[o] = e;
    if (emb16) emb16[o] = __float2bfloat16_rn(e);
  }
}
void launch_embed(const float* img, float* pts, float* emb32, __nv_bfloat16* emb16, int M, int P,
                  const float* dim_t_dev, cudaStream_t st) {
  embed_kernel<<<(M + 3) / 4, 256, 0, st>>>(img, pts, emb32, emb16, M, P, dim_t_dev);
}

// ===================================================================================
// BEV sampling plan, one CTA per scene (GridSampleCrossBEVAttention.forward,
// modules/blocks.py:98-125):
//   aw      = softmax_p(Linear(D->P)(q))                                    (:110-112)
//   grid    = (y / lidar_max_x, x / lidar_max_y)                            (:101-108)
//   ix, iy  = ((g+1)*size-1)/2, 4 bilinear corners, zero padding            (:117-122)
// Output: the sorted list of UNIQUE in-bounds corner pixels of the scene, packed (y << 16) | x
// (the only pixels at which value_proj has to be evaluated), and per (anchor, pose, corner) the slot of its pixel
// in that list with the combined weight bilinear * aw.
// ===================================================================================
__global__ void __launch_bounds__(256, 6) plan_kernel(const float* __restrict__ q0,
                                                   const float* __restrict__ attw_w,
                                                   const float* __restrict__ attw_b,
                                                   const float* __restrict__ pts,
                                                   int* __restrict__ upix, int* __restrict__ nuniq,
                                                   int* __restrict__ ent_slot,
                                                   float* __restrict__ ent_w,
                                                   int* __restrict__ rows_total,
                                                   unsigned int* __restrict__ need_seg,
                                                   unsigned int* __restrict__ done_seg,
                                                   int seg_shift, int nw32,
                                                   int A, int P, int H, int W, int rcap,
                                                   OdoConsts oc, int q0_spt, PlanReuse ru) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int HW = H * W;
  unsigned short* table = reinterpret_cast<unsigned short*>(smraw);        // [HW]
  float* aw = reinterpret_cast<float*>(smraw + ((HW * 2 + 15) / 16) * 16);  // [A*P]
  __shared__ int warp_tot[8];
  __shared__ int total_s;
  __shared__ unsigned int seg_s[64];   // BEV segments (2^seg_shift pixels of a row) the conv will read
  const int scene = blockIdx.x;
  if (threadIdx.x < 64) seg_s[threadIdx.x] = 0u;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // table[pixel] = slot + 1 of a pixel whose value row exists (0: none); bit 15 marks the pixels this
  // call samples.  A later denoise step of the same layer (ru.mode == 2) starts from the table the
  // earlier steps left: value_proj(bev) does not depend on the step, so only pixels without a row
  // are handed to the conv.
  const bool later = ru.mode >= 2;   // a denoise step after the first (3: its pixels all count as new)
  if (ru.mode == 2) {
    const uint4* src = reinterpret_cast<const uint4*>(ru.slot_tab + (size_t)scene * HW);
    for (int i = tid; i < HW / 8; i += 256) reinterpret_cast<uint4*>(table)[i] = __ldg(src + i);
    for (int i = (HW / 8) * 8 + tid; i < HW; i += 256) table[i] = ru.slot_tab[(size_t)scene * HW + i];
  } else {
    for (int i = tid; i < HW / 8; i += 256) reinterpret_cast<uint4*>(table)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = (HW / 8) * 8 + tid; i < HW; i += 256) table[i] = 0;
  }
  __shared__ int reuse_s[2];   // rows the scene already has, base of its rows in the global new-row list
  // ---- attention weights: P (=8) logits per anchor, softmax over the poses; the 8 x 256 head is
  // staged in shared memory once per CTA
  __shared__ __align__(16) float ww[8 * D];
  if (ru.logit_part) {
    // the chain engine's encoder program left the logits as two partial sums per row (kernels_chain.h)
    if (tid < A) {
      const int tile = scene / q0_spt, r = (scene - tile * q0_spt) * A + tid;
      const float* p0 = ru.logit_part + (((size_t)tile * 2) * 128 + r) * ru.logit_ld + ru.logit_off;
      const float* p1 = p0 + (size_t)128 * ru.logit_ld;
      float e[8], mx = -INFINITY, den = 0.f;
#pragma unroll
      for (int o = 0; o < 8; ++o) { e[o] = (p0[o] + p1[o]) + attw_b[o]; mx = fmaxf(mx, e[o]); }
#pragma unroll
      for (int o = 0; o < 8; ++o) { e[o] = expf(e[o] - mx); den += e[o]; }
#pragma unroll
      for (int o = 0; o < 8; ++o) aw[tid * P + o] = e[o] / den;
    }
  } else {
    for (int i = tid; i < 8 * D / 4; i += 256)
      reinterpret_cast<float4*>(ww)[i] = __ldg(reinterpret_cast<const float4*>(attw_w) + i);
    __syncthreads();
  }
  for (int a = warp; a < A && !ru.logit_part; a += 8) {   // (in-kernel logits: other engines, > 16 logits per row)
    float qv[8];
    if (q0_spt > 0) {
      // chain engine layout (kernels_chain.cu): [tile][64 column groups][128 rows] float4
      const int tile = scene / q0_spt, r = (scene - tile * q0_spt) * A + a;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int col = lane + 32 * i;
        qv[i] = q0[(((size_t)tile * 64 + (col >> 2)) * 128 + r) * 4 + (col & 3)];
      }
    } else {
      const float* q = q0 + ((size_t)scene * A + a) * D;
#pragma unroll
      for (int i = 0; i < 8; ++i) qv[i] = q[lane + 32 * i];
    }
    float logit[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s = fmaf(qv[i], ww[o * D + lane + 32 * i], s);
      logit[o] = warp_sum(s) + attw_b[o];
    }
    float mx = logit[0];
#pragma unroll
    for (int o = 1; o < 8; ++o) mx = fmaxf(mx, logit[o]);
    float e[8], den = 0.f;
#pragma unroll
    for (int o = 0; o < 8; ++o) { e[o] = expf(logit[o] - mx); den += e[o]; }
    if (lane < 8) {
      float mine = e[0];
#pragma unroll
      for (int o = 1; o < 8; ++o) if (lane == o) mine = e[o];
      aw[a * P + lane] = mine / den;
    }
  }
  __syncthreads();
  // ---- mark needed pixels
  const int AP = A * P;
  for (int e = tid; e < AP; e += 256) {
    const float px = pts[((size_t)scene * AP + e) * 2 + 0];
    const float py = pts[((size_t)scene * AP + e) * 2 + 1];
    const Corners c = corners_of(px, py, H, W, oc);
#pragma unroll
    for (int k = 0; k < 4; ++k)   // (racing writers of one entry all store the same value)
      if (c.pix[k] >= 0) table[c.pix[k]] = (unsigned short)(table[c.pix[k]] | 0x8000u);
  }
  __syncthreads();
  // ---- ordered compaction of the sampled pixels that have no value row yet (pixel order == memory order of the NHWC map)
  const int ipt = (HW + 255) / 256;
  const int beg = tid * ipt, end = min(HW, beg + ipt);
  int cnt = 0;
  // a thread's run of table entries as a bit mask (runs of 16: two 16-byte reads instead of 16
  // bank-conflicting 2-byte ones); longer runs fall back to the scalar scan
  unsigned int present = 0;
  const bool fast = ipt == 16 && (HW & 15) == 0;
  if (fast) {
    const uint4 t0 = *reinterpret_cast<const uint4*>(table + beg);
    const uint4 t1 = *reinterpret_cast<const uint4*>(table + beg + 8);
    const unsigned int w[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if ((w[k] & 0xffffu) == 0x8000u) present |= 1u << (2 * k);
      if ((w[k] >> 16) == 0x8000u) present |= 1u << (2 * k + 1);
    }
    cnt = __popc(present);
  } else {
    for (int i = beg; i < end; ++i) cnt += table[i] == 0x8000u ? 1 : 0;
  }
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  int base = incl - cnt;
  for (int w = 0; w < warp; ++w) base += warp_tot[w];
  if (tid == 255) {
    const int total = base + cnt;
    total_s = total;
    int have = 0, gbase = 0;
    if (later) {
      have = ru.slot_cnt[scene];
      gbase = total ? atomicAdd(ru.new_count, total) : 0;
    }
    if (ru.mode) ru.slot_cnt[scene] = have + total;
    reuse_s[0] = have;
    reuse_s[1] = gbase;
  }
  if (later) __syncthreads();
  const int have = later ? reuse_s[0] : 0, gbase = later ? reuse_s[1] : 0;
  // segments covering the 3x3 neighbourhood of every unique pixel; a thread's pixels are
  // consecutive, so it marks [xmin - 1, xmax + 1] of rows y - 1 .. y + 1 once per row it touches
  const int segs_per_row = W >> seg_shift;
  int cur_y = -1, xmin = 0, xmax = 0;
  auto flush = [&]() {
    if (cur_y < 0 || !need_seg) return;
    const int s0 = max(xmin - 1, 0) >> seg_shift, s1 = min(xmax + 1, W - 1) >> seg_shift;
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = cur_y + dy;
      if (yy < 0 || yy >= H) continue;
      for (int sg = s0; sg <= s1; ++sg) {
        const int bit = yy * segs_per_row + sg;
        atomicOr(&seg_s[bit >> 5], 1u << (bit & 31));
      }
    }
  };
  auto take = [&](int i) {
    const int yy = i / W, xx = i - yy * W;
    if (later) {   // row of the cross-scene list: (pixel of the batch, value row it fills)
      DDH_ASSERT(have + base < ru.vcap);
      ru.new_list[gbase + base] = make_int2(scene * HW + i, scene * ru.vcap + have + base);
    } else {
      DDH_ASSERT(base < rcap);
      upix[(size_t)scene * rcap + base] = (yy << 16) | xx;   // packed (y, x)
    }
    table[i] = (unsigned short)(have + base + 1);
    ++base;
    if (yy != cur_y) { flush(); cur_y = yy; xmin = xx; }
    xmax = xx;
  };
  if (fast) {   // walk the set bits only (a thread's 16 pixels hold ~1 sampled pixel on average)
    for (unsigned int m = present; m; m &= m - 1) take(beg + __ffs((int)m) - 1);
  } else {
    for (int i = beg; i < end; ++i)
      if (table[i] == 0x8000u) take(i);
  }
  flush();
  __syncthreads();
  if (tid == 0) {
    if (!later) nuniq[scene] = total_s;
    if (rows_total) atomicAdd(rows_total, total_s);
  }
  if (ru.mode && ru.keep) {   // slot table for the next denoise step
    uint4* dst = reinterpret_cast<uint4*>(ru.slot_tab + (size_t)scene * HW);
    for (int i = tid; i < HW / 8; i += 256) {
      uint4 t = reinterpret_cast<const uint4*>(table)[i];
      t.x &= 0x7fff7fffu; t.y &= 0x7fff7fffu; t.z &= 0x7fff7fffu; t.w &= 0x7fff7fffu;
      dst[i] = t;
    }
    for (int i = (HW / 8) * 8 + tid; i < HW; i += 256)
      ru.slot_tab[(size_t)scene * HW + i] = (unsigned short)(table[i] & 0x7fffu);
  }
  if (need_seg && tid < nw32) {   // segments still to convert for the coming conv call; mark them converted
    const unsigned int need = seg_s[tid];
    const unsigned int done = done_seg[(size_t)scene * nw32 + tid];
    need_seg[(size_t)scene * nw32 + tid] = need & ~done;
    done_seg[(size_t)scene * nw32 + tid] = done | need;
  }
  // ---- entries
  for (int e = tid; e < AP; e += 256) {
    const float px = pts[((size_t)scene * AP + e) * 2 + 0];
    const float py = pts[((size_t)scene * AP + e) * 2 + 1];
    const Corners c = corners_of(px, py, H, W, oc);
    const float a_w = aw[e];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const size_t o = ((size_t)scene * AP + e) * 4 + k;
      if (c.pix[k] >= 0) {
        ent_slot[o] = (int)(table[c.pix[k]] & 0x7fffu) - 1;
        ent_w[o] = c.w[k] * a_w;
      } else {
        ent_slot[o] = -1;
        ent_w[o] = 0.f;
      }
    }
  }
}
void launch_plan(const float* q0, const float* attw_w, const float* attw_b, const float* pts,
                 int* upix, int* nuniq, int* ent_slot, float* ent_w, int* rows_total,
                 unsigned int* need_seg, unsigned int* done_seg, int seg_shift, int nw32, int B, int A,
                 int P, int H, int W, int rcap, OdoConsts oc, cudaStream_t st, int q0_spt,
                 PlanReuse ru) {
  const int smem = ((H * W * 2 + 15) / 16) * 16 + A * P * 4;
  static int cur = 0;
  if (smem > cur) {
    cudaFuncSetAttribute(plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cur = smem;
  }
  plan_kernel<<<B, 256, smem, st>>>(q0, attw_w, attw_b, pts, upix, nuniq, ent_slot, ent_w,
                                     rows_total, need_seg, done_seg, seg_shift, nw32, A, P, H, W,
                                     rcap, oc, q0_spt, ru);
}

// ===================================================================================
// Bilinear + attention-weighted combine (modules/blocks.py:117-126):
//   S[a, c] = sum_{p, corner} (bilinear * aw)[a,p,corner] * value[pixel(a,p,corner), c]
// One CTA per scene, one thread per channel; value rows were produced at the unique pixels.
// ===================================================================================
__global__ void __launch_bounds__(256) combine_kernel(const float* __restrict__ V,
                                                      const int* __restrict__ ent_slot,
                                                      const float* __restrict__ ent_w,
                                                      float* __restrict__ s32,
                                                      __nv_bfloat16* __restrict__ s16, int A, int P,
                                                      int rcap) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int scene = blockIdx.x, c = threadIdx.x;
  const int n_ent = A * P * 4;
  int* sl = reinterpret_cast<int*>(smraw);
  float* sw = reinterpret_cast<float*>(smraw + (size_t)n_ent * 4);
  for (int i = c; i < n_ent; i += 256) {
    sl[i] = ent_slot[(size_t)scene * n_ent + i];
    sw[i] = ent_w[(size_t)scene * n_ent + i];
  }
  __syncthreads();
  const float* Vs = V + (size_t)scene * rcap * D + c;
  const int per = P * 4;
  for (int a = 0; a < A; ++a) {
    float acc = 0.f;
#pragma unroll 8
    for (int j = 0; j < per; ++j) {
      const int s = sl[a * per + j];
      if (s >= 0) acc = fmaf(sw[a * per + j], Vs[(size_t)s * D], acc);
    }
    const size_t o = ((size_t)scene * A + a) * D + c;
    if (s32) s32[o] = acc;
    if (s16) s16[o] = __float2bfloat16_rn(acc);
  }
}
void launch_combine(const float* V, const int* ent_slot, const float* ent_w, float* s32,
                    __nv_bfloat16* s16, int B, int A, int P, int rcap, cudaStream_t st) {
  const int smem = A * P * 4 * 8;
  static int cur = 0;
  if (smem > cur && smem > 48 * 1024) {
    cudaFuncSetAttribute(combine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cur = smem;
  }
  combine_kernel<<<B, 256, smem, st>>>(V, ent_slot, ent_w, s32, s16, A, P, rcap);
}

// ===================================================================================
// The same combine over the kept value rows of a layer (denoise steps after the first, PlanReuse):
//   S[a, :] = sum_k w[a, k] * V[scene * vcap + slot[a, k], :]      k = (pose, corner), in entry order
// One CTA per scene, one warp per anchor (round robin), one lane per 8 channels: an entry is one 512-byte
// row read by the whole warp (L1/L2 hits after the scene's first touch of a row); lane k holds entry k.
// ===================================================================================
__global__ void __launch_bounds__(256, 4) combine_rows_kernel(const __nv_bfloat16* __restrict__ V,
                                                           const int* __restrict__ ent_slot,
                                                           const float* __restrict__ ent_w,
                                                           __nv_bfloat16* __restrict__ s16, int A,
                                                           int epa, int vcap) {
  const int scene = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint4* rows = reinterpret_cast<const uint4*>(V) + (size_t)scene * vcap * 32 + lane;
  for (int a = warp; a < A; a += 8) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    const size_t e0 = ((size_t)scene * A + a) * epa;
    for (int kb = 0; kb < epa; kb += 32) {
      const int n = min(32, epa - kb);
      int slot = -1;
      float wv = 0.f;
      if (lane < n) { slot = __ldg(ent_slot + e0 + kb + lane); wv = __ldg(ent_w + e0 + kb + lane); }
      DDH_ASSERT(slot < vcap);
#pragma unroll 1
      for (int k0 = 0; k0 < n; k0 += 8) {
        uint4 v[8];
        float wk[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int sl = __shfl_sync(0xffffffffu, slot, (k0 + u) & 31);
          wk[u] = __shfl_sync(0xffffffffu, wv, (k0 + u) & 31);
          const bool ok = sl >= 0 && k0 + u < n;   // warp-uniform; a missing corner reads nothing
          v[u] = ok ? __ldg(rows + (size_t)sl * 32) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t w4[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
          for (int h2 = 0; h2 < 4; ++h2) {
            acc[2 * h2] = fmaf(wk[u], __uint_as_float(w4[h2] << 16), acc[2 * h2]);
            acc[2 * h2 + 1] = fmaf(wk[u], __uint_as_float(w4[h2] & 0xffff0000u), acc[2 * h2 + 1]);
          }
        }
      }
    }
    uint4 o;
    uint32_t* op = &o.x;
#pragma unroll
    for (int h2 = 0; h2 < 4; ++h2) {
      const __nv_bfloat162 hh = __floats2bfloat162_rn(acc[2 * h2], acc[2 * h2 + 1]);
      op[h2] = *reinterpret_cast<const uint32_t*>(&hh);
    }
    reinterpret_cast<uint4*>(s16)[((size_t)scene * A + a) * 32 + lane] = o;
  }
}
void launch_combine_rows(const __nv_bfloat16* V, const int* ent_slot, const float* ent_w,
                         __nv_bfloat16* s16, int B, int A, int ent_per_anchor, int vcap, cudaStream_t st) {
  combine_rows_kernel<<<B, 256, 0, st>>>(V, ent_slot, ent_w, s16, A, ent_per_anchor, vcap);
}

// ===================================================================================
// Agent cross-attention core (nn.MultiheadAttention, transfuser_model_v2.py:316-321,355-357)
// after the Q projection and the hoisted K|V projection: per scene and head,
// softmax(q*scale . K^T) . V with head_dim 32 and Na <= 32 keys.
// One CTA per scene, one warp per head, one LANE per query: the scene's K|V rows sit in shared
// memory and are read as warp-wide broadcasts, the query / score / output vectors live in
// registers, so the inner loops are pure LDS.128 + FFMA (no shuffles).
// ===================================================================================
__global__ void __launch_bounds__(256, 2) attn_core_kernel(const float* __restrict__ qh,
                                                           const float* __restrict__ kv,
                                                           float* __restrict__ o32,
                                                           __nv_bfloat16* __restrict__ o16, int A,
                                                           int Na, int heads) {
  extern __shared__ __align__(16) float kv_s[];   // [Na][2*D]: K | V
  const int scene = blockIdx.x;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    // 8 independent 16-byte loads in flight per thread (a load-then-store loop would serialise
    // one memory round trip per iteration: 15 of them for 60 KB of K|V)
    const float4* src = reinterpret_cast<const float4*>(kv + (size_t)scene * Na * 2 * D);
    float4* dst = reinterpret_cast<float4*>(kv_s);
    const int n4 = Na * 2 * D / 4;
    for (int base = threadIdx.x; base < n4; base += 256 * 8) {
      float4 t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) if (base + u * 256 < n4) t[u] = __ldg(src + base + u * 256);
#pragma unroll
      for (int u = 0; u < 8; ++u) if (base + u * 256 < n4) dst[base + u * 256] = t[u];
    }
  }
  __syncthreads();
  if (h >= heads) return;
  const float scale = 0.17677669529663687f;  // 1/sqrt(32)
  for (int a0 = 0; a0 < A; a0 += 32) {
    const int a = a0 + lane;
    const bool act = a < A;
    const size_t row = (size_t)scene * A + (act ? a : 0);
    float q[32];
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(qh + row * D + h * 32) + c4);
      q[4 * c4 + 0] = t.x * scale; q[4 * c4 + 1] = t.y * scale;
      q[4 * c4 + 2] = t.z * scale; q[4 * c4 + 3] = t.w * scale;
    }
    float s[32];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      s[j] = -INFINITY;
      if (j < Na) {
        const float4* kr = reinterpret_cast<const float4*>(kv_s + j * 2 * D + h * 32);
        float acc = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 k4 = kr[c4];
          acc = fmaf(q[4 * c4 + 0], k4.x, acc);
          acc = fmaf(q[4 * c4 + 1], k4.y, acc);
          acc = fmaf(q[4 * c4 + 2], k4.z, acc);
          acc = fmaf(q[4 * c4 + 3], k4.w, acc);
        }
        s[j] = acc;
        mx = fmaxf(mx, acc);
      }
    }
    float den = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      s[j] = (j < Na) ? expf(s[j] - mx) : 0.f;
      den += s[j];
    }
    const float inv = 1.0f / den;
    float o[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) o[c] = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j < Na) {
        const float4* vr = reinterpret_cast<const float4*>(kv_s + j * 2 * D + D + h * 32);
        const float pj = s[j] * inv;
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 v4 = vr[c4];
          o[4 * c4 + 0] = fmaf(pj, v4.x, o[4 * c4 + 0]);
          o[4 * c4 + 1] = fmaf(pj, v4.y, o[4 * c4 + 1]);
          o[4 * c4 + 2] = fmaf(pj, v4.z, o[4 * c4 + 2]);
          o[4 * c4 + 3] = fmaf(pj, v4.w, o[4 * c4 + 3]);
        }
      }
    }
    if (act) {
      if (o32) {
        float4* d = reinterpret_cast<float4*>(o32 + row * D + h * 32);
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4)
          d[c4] = make_float4(o[4 * c4], o[4 * c4 + 1], o[4 * c4 + 2], o[4 * c4 + 3]);
      }
      if (o16) {
        uint4* d = reinterpret_cast<uint4*>(o16 + row * D + h * 32);
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          __nv_bfloat162 h0 = __floats2bfloat162_rn(o[8 * c8 + 0], o[8 * c8 + 1]);
          __nv_bfloat162 h1 = __floats2bfloat162_rn(o[8 * c8 + 2], o[8 * c8 + 3]);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(o[8 * c8 + 4], o[8 * c8 + 5]);
          __nv_bfloat162 h3 = __floats2bfloat162_rn(o[8 * c8 + 6], o[8 * c8 + 7]);
          uint4 u;
          u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
          u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
          d[c8] = u;
        }
      }
    }
  }
}
void launch_attn_core(const float* qh, const float* kv, float* o32, __nv_bfloat16* o16, int B,
                      int A, int Na, int heads, cudaStream_t st) {
  const int smem = Na * 2 * D * 4;
  static int cur = 0;
  if (smem > cur) {
    cudaFuncSetAttribute(attn_core_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cur = smem;
  }
  attn_core_kernel<<<B, 256, smem, st>>>(qh, kv, o32, o16, A, Na, heads);
}

// ===================================================================================
// Regression tail of DiffMotionPlanningRefinementModule + layer/step bookkeeping:
//   reg = Linear(D->3P)(r2)                         (transfuser_model_v2.py:225-231,253-254)
//   reg[..., :2] += points ; reg[..., 2] = tanh(.)*pi                        (:378-380)
//   next layer's points = reg[..., :2]                                       (:424)
//   after the last layer of a non-final step: x0 = norm_odo(reg[..., :2]);
//   img = DDIM step(x0, t, img), prediction_type "sample", clip_sample, eta 0 (:632-636)
// One warp per anchor row.
// ===================================================================================
__global__ void __launch_bounds__(256) reg_finish_kernel(const float* __restrict__ r2,
                                                         const float* __restrict__ w4,
                                                         const float* __restrict__ b4,
                                                         float* __restrict__ pts,
                                                         float* __restrict__ img,
                                                         float* __restrict__ modes, int M, int P,
                                                         int do_ddim, DdimCoef dc) {
  // 4 lanes per row (interleaved float4 columns), 8 rows per warp, 64 rows per CTA; the
  // [24][256] head weights are read from shared memory as (4-address) broadcasts.
  __shared__ __align__(16) float w_s[24 * D];
  __shared__ float b_s[24];
  {
    float4 t[6];   // 24*256/4 = 1536 float4 = 6 per thread, all in flight
#pragma unroll
    for (int u = 0; u < 6; ++u) t[u] = __ldg(reinterpret_cast<const float4*>(w4) + threadIdx.x + u * 256);
#pragma unroll
    for (int u = 0; u < 6; ++u) reinterpret_cast<float4*>(w_s)[threadIdx.x + u * 256] = t[u];
  }
  if (threadIdx.x < 24) b_s[threadIdx.x] = b4[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int part = lane & 3;
  const int m = blockIdx.x * 64 + warp * 8 + (lane >> 2);
  const bool valid = m < M;
  const float* row = r2 + (size_t)(valid ? m : 0) * D;
  float acc[24];
#pragma unroll
  for (int o = 0; o < 24; ++o) acc[o] = 0.f;
#pragma unroll 4
  for (int k = 0; k < 16; ++k) {
    const int f = k * 4 + part;                       // float4 index within the row
    const float4 r = __ldg(reinterpret_cast<const float4*>(row) + f);
#pragma unroll
    for (int o = 0; o < 24; ++o) {
      const float4 w = reinterpret_cast<const float4*>(w_s + o * D)[f];
      acc[o] = fmaf(r.x, w.x, acc[o]);
      acc[o] = fmaf(r.y, w.y, acc[o]);
      acc[o] = fmaf(r.z, w.z, acc[o]);
      acc[o] = fmaf(r.w, w.w, acc[o]);
    }
  }
#pragma unroll
  for (int o = 0; o < 24; ++o) {
    acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], 1);
    acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], 2);
  }
  if (!valid) return;
  // lane `part` finishes outputs part*6 .. part*6+5  (= poses 2*part, 2*part+1)
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    const int o = part * 6 + q;
    float mine = 0.f;
#pragma unroll
    for (int oo = 0; oo < 24; ++oo) if (oo == o) mine = acc[oo];
    mine += b_s[o];
    const int p = o / 3, comp = o - p * 3;
    float out;
    if (comp < 2) {
      const size_t pi = ((size_t)m * P + p) * 2 + comp;
      out = __fadd_rn(mine, pts[pi]);
      pts[pi] = out;
      if (do_ddim) {
        const float x0 = comp ? norm_y(out) : norm_x(out);
        const float sample = img[pi];
        const float eps = __fdiv_rn(__fsub_rn(sample, __fmul_rn(dc.sqrt_ac_t, x0)), dc.sqrt_1m_ac_t);
        const float x0c = fminf(fmaxf(x0, -1.0f), 1.0f);
        img[pi] = __fadd_rn(__fmul_rn(dc.sqrt_ac_prev, x0c), __fmul_rn(dc.sqrt_1m_ac_prev, eps));
      }
    } else {
      out = __fmul_rn(tanhf(mine), 3.14159265358979323846f);
    }
    modes[((size_t)m * P + p) * 3 + comp] = out;
  }
}
void launch_reg_finish(const float* r2, const float* w4, const float* b4, float* pts, float* img,
                       float* modes, int M, int P, int do_ddim, DdimCoef dc, cudaStream_t st) {
  reg_finish_kernel<<<(M + 63) / 64, 256, 0, st>>>(r2, w4, b4, pts, img, modes, M, P, do_ddim, dc);
}

// ===================================================================================
// Generic small multi-head attention (query decoder of V2TransfuserModel, nn.TransformerDecoder
// self- and cross-attention, transfuser_model_v2.py:73-80,141-146): per scene and head
// softmax(q k^T / sqrt(32)) v with head_dim 32, up to 32 queries and any number of keys that fits
// shared memory.  One CTA per scene, one warp per head, one lane per query; K|V rows are staged in
// shared memory (fp32) and read as broadcasts; online softmax over the keys.
//   q  [B*nq rows][ldq],   kv [B*nk rows][ldkv] with K at column 0 and V at column 256 of the row
// ===================================================================================
__global__ void __launch_bounds__(256) mha_small_kernel(const float* __restrict__ q, int ldq,
                                                        const float* __restrict__ kv, int ldkv,
                                                        float* __restrict__ o32,
                                                        __nv_bfloat16* __restrict__ o16, int nq, int nk) {
  extern __shared__ __align__(16) float kvs[];   // [nk][512]
  const int scene = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < nk * 128; i += 256) {
    const int r = i >> 7, c4 = i & 127;
    reinterpret_cast<float4*>(kvs)[i] = __ldg(reinterpret_cast<const float4*>(kv + ((size_t)scene * nk + r) * ldkv) + c4);
  }
  __syncthreads();
  const bool act = lane < nq;
  const size_t row = (size_t)scene * nq + (act ? lane : 0);
  const float scale = 0.17677669529663687f;   // 1 / sqrt(32)
  float qv[32], o[32];
#pragma unroll
  for (int c4 = 0; c4 < 8; ++c4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(q + row * ldq + h * 32) + c4);
    qv[4 * c4] = t.x * scale; qv[4 * c4 + 1] = t.y * scale; qv[4 * c4 + 2] = t.z * scale; qv[4 * c4 + 3] = t.w * scale;
  }
#pragma unroll
  for (int c = 0; c < 32; ++c) o[c] = 0.f;
  float mx = -INFINITY, den = 0.f;
  for (int j = 0; j < nk; ++j) {
    const float4* kr = reinterpret_cast<const float4*>(kvs + (size_t)j * 512 + h * 32);
    float sc = 0.f;
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      const float4 k4 = kr[c4];
      sc = fmaf(qv[4 * c4], k4.x, sc); sc = fmaf(qv[4 * c4 + 1], k4.y, sc);
      sc = fmaf(qv[4 * c4 + 2], k4.z, sc); sc = fmaf(qv[4 * c4 + 3], k4.w, sc);
    }
    const float nm = fmaxf(mx, sc);
    const float alpha = expf(mx - nm), pj = expf(sc - nm);
    den = den * alpha + pj;
    mx = nm;
    const float4* vr = reinterpret_cast<const float4*>(kvs + (size_t)j * 512 + 256 + h * 32);
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      const float4 v4 = vr[c4];
      o[4 * c4] = fmaf(pj, v4.x, o[4 * c4] * alpha); o[4 * c4 + 1] = fmaf(pj, v4.y, o[4 * c4 + 1] * alpha);
      o[4 * c4 + 2] = fmaf(pj, v4.z, o[4 * c4 + 2] * alpha); o[4 * c4 + 3] = fmaf(pj, v4.w, o[4 * c4 + 3] * alpha);
    }
  }
  if (!act) return;
  const float inv = 1.0f / den;
  if (o32) {
    float4* d = reinterpret_cast<float4*>(o32 + row * D + h * 32);
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4)
      d[c4] = make_float4(o[4 * c4] * inv, o[4 * c4 + 1] * inv, o[4 * c4 + 2] * inv, o[4 * c4 + 3] * inv);
  }
  if (o16) {
    uint4* d = reinterpret_cast<uint4*>(o16 + row * D + h * 32);
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
      __nv_bfloat162 h0 = __floats2bfloat162_rn(o[8 * c8] * inv, o[8 * c8 + 1] * inv);
      __nv_bfloat162 h1 = __floats2bfloat162_rn(o[8 * c8 + 2] * inv, o[8 * c8 + 3] * inv);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(o[8 * c8 + 4] * inv, o[8 * c8 + 5] * inv);
      __nv_bfloat162 h3 = __floats2bfloat162_rn(o[8 * c8 + 6] * inv, o[8 * c8 + 7] * inv);
      uint4 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
      u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
      d[c8] = u;
    }
  }
}
int launch_mha_small(const float* q, int ldq, const float* kv, int ldkv, float* o32, __nv_bfloat16* o16,
                     int B, int nq, int nk, cudaStream_t st) {
  const int smem = nk * 512 * 4;
  static int cur = 0;
  if (smem > cur) {
    cudaError_t e = cudaFuncSetAttribute(mha_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    cur = smem;
  }
  mha_small_kernel<<<B, 256, smem, st>>>(q, ldq, kv, ldkv, o32, o16, nq, nk);
  return 0;
}

// y[m][o] = x[m] . w[o] + b[o] for a handful of outputs per row (AgentHead's last Linears,
// transfuser_model_v2.py:187-203), one warp per row.  agent_states: tanh * 32 on outputs 0-1,
// tanh * pi on output 2.
__global__ void __launch_bounds__(256) rowdot_kernel(const float* __restrict__ x, int ldx,
                                                     const float* __restrict__ w, const float* __restrict__ b,
                                                     float* __restrict__ y, int M, int K, int n_out,
                                                     int rows_per_group, int skip_first, int states) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int out_rows_per_group = rows_per_group - skip_first;
  if (r >= M) return;
  const int g = r / rows_per_group, i = r - g * rows_per_group;
  if (i < skip_first) return;
  const float* xr = x + (size_t)r * ldx;
  for (int o = 0; o < n_out; ++o) {
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s = fmaf(xr[k], w[(size_t)o * K + k], s);
    s = warp_sum(s) + b[o];
    if (states) s = (o < 2) ? tanhf(s) * 32.0f : (o == 2 ? tanhf(s) * 3.14159265358979323846f : s);
    if (lane == 0) y[((size_t)g * out_rows_per_group + (i - skip_first)) * n_out + o] = s;
  }
}
void launch_rowdot(const float* x, int ldx, const float* w, const float* b, float* y, int M, int K, int n_out,
                   int rows_per_group, int skip_first, int states, cudaStream_t st) {
  rowdot_kernel<<<(M + 7) / 8, 256, 0, st>>>(x, ldx, w, b, y, M, K, n_out, rows_per_group, skip_first, states);
}

// x[b][q][:] = emb[q][:]  (the query embedding is the first layer's input for every scene, :141)
__global__ void broadcast_rows_kernel(const float* __restrict__ emb, float* __restrict__ x32,
                                      __nv_bfloat16* __restrict__ x16, int rows_per_group, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / D;
    const float v = emb[(r % rows_per_group) * D + (i - r * D)];
    x32[i] = v;
    if (x16) x16[i] = __float2bfloat16_rn(v);
  }
}
void launch_broadcast_rows(const float* emb, float* x32, __nv_bfloat16* x16, int rows_per_group, size_t n,
                           cudaStream_t st) {
  const int blocks = (int)min((size_t)148 * 8, (n + 255) / 256);
  broadcast_rows_kernel<<<blocks, 256, 0, st>>>(emb, x32, x16, rows_per_group, n);
}

// mode = argmax(cls) (first maximum wins), trajectory = reg[b, mode]          (:637-640)
__global__ void select_kernel(const float* __restrict__ scores, const float* __restrict__ modes,
                              float* __restrict__ traj, long long* __restrict__ mode_idx, int B,
                              int A, int P) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int best = 0;
  float bv = scores[(size_t)b * A];
  for (int a = 1; a < A; ++a) {
    const float v = scores[(size_t)b * A + a];
    if (v > bv) { bv = v; best = a; }
  }
  if (mode_idx) mode_idx[b] = best;
  if (traj)
    for (int i = 0; i < P * 3; ++i) traj[(size_t)b * P * 3 + i] = modes[((size_t)b * A + best) * P * 3 + i];
}
void launch_select(const float* scores, const float* modes, float* traj, long long* mode_idx,
                   int B, int A, int P, cudaStream_t st) {
  select_kernel<<<(B + 127) / 128, 128, 0, st>>>(scores, modes, traj, mode_idx, B, A, P);
}

// ===================================================================================
// Pack-time helpers (run once per ddh_pack_weights).
// ===================================================================================
__global__ void transpose_f32_kernel(const float* __restrict__ s, float* __restrict__ d, int rows,
                                     int cols) {
  __shared__ float t[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    t[i][threadIdx.x] = (r < rows && c < cols) ? s[(size_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) d[(size_t)c * rows + r] = t[threadIdx.x][i];
  }
}
void launch_transpose_f32(const float* src, float* dst, int rows, int cols, cudaStream_t st) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_f32_kernel<<<grid, block, 0, st>>>(src, dst, rows, cols);
}

// conv weight [Cout][Cin][3][3]  ->  Wt[(tap*Cin + c)][Cout]  (fp32, k-major for the SIMT engine)
__global__ void pack_conv_f32_kernel(const float* __restrict__ w, float* __restrict__ d, int Cout,
                                     int Cin) {
  const size_t n = (size_t)Cout * Cin * 9;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const int o = (int)(i % Cout);
    const size_t k = i / Cout;
    const int c = (int)(k % Cin), tap = (int)(k / Cin);
    d[i] = w[((size_t)o * Cin + c) * 9 + tap];
  }
}
// conv weight -> W[Cout][(tap*Cin + c)] bf16 (K contiguous, the tensor engine's B operand)
__global__ void pack_conv_bf16_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ d,
                                      int Cout, int Cin) {
  const size_t n = (size_t)Cout * Cin * 9;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t k = i % ((size_t)Cin * 9);
    const int o = (int)(i / ((size_t)Cin * 9));
    const int c = (int)(k % Cin), tap = (int)(k / Cin);
    d[i] = __float2bfloat16_rn(w[((size_t)o * Cin + c) * 9 + tap]);
  }
}
// conv weight -> fp32 [Cout][w_hi (9*Cin) | w_lo (9*Cin)], K ordered (tap, channel): the B operand of the
// 3xTF32 conv (w_hi = the 10 mantissa bits the tensor core reads, w_lo = w - w_hi, exact)
__global__ void pack_conv_tf32x2_kernel(const float* __restrict__ w, float* __restrict__ d, int Cout, int Cin) {
  const size_t K = (size_t)Cin * 9, n = (size_t)Cout * K;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t k = i % K;
    const int o = (int)(i / K);
    const int c = (int)(k % Cin), tap = (int)(k / Cin);
    const float x = w[((size_t)o * Cin + c) * 9 + tap];
    const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    d[(size_t)o * 2 * K + k] = hi;
    d[(size_t)o * 2 * K + K + k] = x - hi;
  }
}
void launch_pack_conv_tf32x2(const float* w, float* dst, int Cout, int Cin, cudaStream_t st) {
  pack_conv_tf32x2_kernel<<<592, 256, 0, st>>>(w, dst, Cout, Cin);
}
void launch_pack_conv_f32(const float* w, float* dst, int Cout, int Cin, cudaStream_t st) {
  pack_conv_f32_kernel<<<592, 256, 0, st>>>(w, dst, Cout, Cin);
}
void launch_pack_conv_bf16(const float* w, __nv_bfloat16* dst, int Cout, int Cin,
                           cudaStream_t st) {
  pack_conv_bf16_kernel<<<592, 256, 0, st>>>(w, dst, Cout, Cin);
}

// reg head weight [n_out][k] fp32 -> bf16 [64][k]: rows 0.. hold the bf16 rounding (hi) of W, rows
// 32.. the bf16 rounding of the remainder (lo), so that r . (hi + lo) carries ~16 mantissa bits of W
__global__ void pack_hilo_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ d, int n_out,
                                 int k) {
  const int n = 64 * k;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int row = i / k, col = i - row * k;
    const int o = row & 31;
    float v = 0.f;
    if (o < n_out) {
      const float x = w[(size_t)o * k + col];
      const float hi = __bfloat162float(__float2bfloat16_rn(x));
      v = (row < 32) ? hi : (x - hi);
    }
    d[i] = __float2bfloat16_rn(v);
  }
}
void launch_pack_hilo(const float* w, __nv_bfloat16* dst, int n_out, int k, cudaStream_t st) {
  pack_hilo_kernel<<<64, 256, 0, st>>>(w, dst, n_out, k);
}

// y[o] = sum_k W[o][k] * act(x[k]) + b[o]   (torch Linear layout), one warp per output
__global__ void matvec_kernel(const float* __restrict__ W, const float* __restrict__ x,
                              const float* __restrict__ b, float* __restrict__ y, int n_out, int k,
                              int act_in_mish) {
  const int o = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (o >= n_out) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int i = lane; i < k; i += 32) {
    float xv = x[i];
    if (act_in_mish) xv = mishf(xv);
    s = fmaf(W[(size_t)o * k + i], xv, s);
  }
  s = warp_sum(s);
  if (lane == 0) y[o] = s + (b ? b[o] : 0.f);
}
void launch_matvec(const float* W, const float* x, const float* b, float* y, int n_out, int k,
                   int act_in_mish, cudaStream_t st) {
  matvec_kernel<<<(n_out + 7) / 8, 256, 0, st>>>(W, x, b, y, n_out, k, act_in_mish);
}

// SinusoidalPosEmb(dim) of one integer timestep (modules/conditional_unet1d.py:53-66)
__global__ void time_sinemb_kernel(float* emb, int dim, int timestep, float c) {
  const int i = threadIdx.x;
  const int half = dim / 2;
  if (i >= half) return;
  const float f = expf(__fmul_rn((float)i, c));
  const float a = __fmul_rn((float)timestep, f);
  emb[i] = sinf(a);
  emb[half + i] = cosf(a);
}
void launch_time_sinemb(float* emb, int dim, int timestep, cudaStream_t st) {
  // python: emb = math.log(10000) / (half_dim - 1) in double, then cast by torch to fp32
  const float c = (float)(-(log(10000.0) / (double)(dim / 2 - 1)));
  time_sinemb_kernel<<<1, dim / 2, 0, st>>>(emb, dim, timestep, c);
}

}  // namespace ddh
